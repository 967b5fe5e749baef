// spectre_long.cu -- pre/post passes of the long-context (n_fft = 8192, 16384) two-pass path.
//
// n_total = R * 4096 (R = 2, 4).  One decimation-in-frequency radix-R stage over rows u + 4096 m turns the long
// transform into R interleaved 4096-point sub-transforms:
//     a_q[u]      = W_N^{u q} * sum_m z[u + 4096 m] W_R^{m q}            (pre pass, this file)
//     Y[q + R k'] = Gfull[q + R k'] * FFT_4096(a_q)[k'] (+ memory)        (the shared-memory kernel, Plan<...,SUB>)
//     b_q         = IFFT_4096(Y[q + R .]) / N
//     y[u+4096 m] = sum_q W_R^{-m q} W_N^{-u q} b_q[u]                    (post pass, this file)
// z = v[c] + i v[c+2] (and v[c+1] + i v[c+3]) is the same two-real-channels-per-complex-lane packing as everywhere
// else, so the intermediate tensor [B][R][4096][C] has exactly the bytes of V: each pass streams it once.
// Both passes are pure streaming kernels (one 16-byte element per thread and row, fully coalesced).
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "spectre_mix_kernel.cuh"
#include "spectre_long.h"

namespace spx {
namespace {

template <class TIO>
__device__ __forceinline__ Cx<float2> load_quad(const TIO *p);
template <>
__device__ __forceinline__ Cx<float2> load_quad<float>(const float *p) {
    const float4 f = __ldcs(reinterpret_cast<const float4 *>(p));
    return {make_float2(f.x, f.y), make_float2(f.z, f.w)};
}
template <>
__device__ __forceinline__ Cx<float2> load_quad<__nv_bfloat16>(const __nv_bfloat16 *p) {
    return GIO<MODE_QUAD, __nv_bfloat16>::load(p);
}
template <class TIO>
__device__ __forceinline__ void store_quad(TIO *p, const Cx<float2> &c) { GIO<MODE_QUAD, TIO>::store(p, c); }

// PRE: user V -> scratch;  POST: scratch -> user out.  sub = 4096 rows per sub-transform.
// twid: W_N^{u q} for q = 1 .. R-1, u < sub, built on the host in double ([q-1][u]; 64 KB / 96 KB for R = 2 / 4, L2-resident:
// every thread of a row reads the same R-1 entries)
template <int R, class TIO, bool PRE>
__global__ void __launch_bounds__(256) long_pass_kernel(const void *src_, void *dst_, long long u_sb, long long u_sn, int B, int rows,
                                                        int C, int sub, const float2 *__restrict__ twid) {
    const int CQ = C / 4;
    const long long items = (long long)B * sub * CQ;
    for (long long it = (long long)blockIdx.x * blockDim.x + threadIdx.x; it < items; it += (long long)gridDim.x * blockDim.x) {
        const int cq = (int)(it % CQ);
        const long long r = it / CQ;
        const int u = (int)(r % sub);
        const int b = (int)(r / sub);
        Cx<float2> x[R];
        if (PRE) {
            const TIO *src = reinterpret_cast<const TIO *>(src_) + (long long)b * u_sb + cq * 4;
#pragma unroll
            for (int m = 0; m < R; ++m) {
                const int row = u + sub * m;
                if (row < rows) x[m] = load_quad<TIO>(src + (long long)row * u_sn);
                else x[m] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
            }
            Dft<R, float2>::run(x);
        } else {
            const float *src = reinterpret_cast<const float *>(src_) + ((long long)b * R * sub + u) * C + cq * 4;
#pragma unroll
            for (int q = 0; q < R; ++q) x[q] = cswap(load_quad<float>(src + (long long)q * sub * C));
        }
        // twiddles W_N^{u q} from the table; the inverse side works on (im, re)-swapped data, where multiplying by W is
        // multiplying the true value by conj(W)
#pragma unroll
        for (int q = 1; q < R; ++q) {
            const float2 w = __ldg(twid + (size_t)(q - 1) * sub + u);
            x[q] = cmul(x[q], w.x, w.y);
        }
        if (PRE) {
            float *dst = reinterpret_cast<float *>(dst_) + ((long long)b * R * sub + u) * C + cq * 4;
#pragma unroll
            for (int q = 0; q < R; ++q) store_quad<float>(dst + (long long)q * sub * C, x[q]);
        } else {
            Dft<R, float2>::run(x);
            TIO *dst = reinterpret_cast<TIO *>(dst_) + (long long)b * u_sb + cq * 4;
#pragma unroll
            for (int m = 0; m < R; ++m) {
                const int row = u + sub * m;
                if (row < rows) store_quad<TIO>(dst + (long long)row * u_sn, cswap(x[m]));
            }
        }
    }
}

template <int R, class TIO, bool PRE>
cudaError_t launch_one(const void *src, void *dst, long long u_sb, long long u_sn, int B, int rows, int C, int sub, int sms,
                       const float2 *twid, cudaStream_t st) {
    const long long items = (long long)B * sub * (C / 4);
    long long blocks = (items + 255) / 256;
    const long long cap = (long long)sms * 16;
    if (blocks > cap) blocks = cap;
    long_pass_kernel<R, TIO, PRE><<<(int)blocks, 256, 0, st>>>(src, dst, u_sb, u_sn, B, rows, C, sub, twid);
    return cudaGetLastError();
}

}  // namespace

cudaError_t long_pass(bool pre, int R, int dtype_bf16, const void *src, void *dst, long long u_sb, long long u_sn, int B, int rows,
                      int C, int sub, int sms, const float2 *twid, cudaStream_t st) {
#define SPX_LP(RR, T, P) return launch_one<RR, T, P>(src, dst, u_sb, u_sn, B, rows, C, sub, sms, twid, st)
    if (R == 2) {
        if (dtype_bf16) { if (pre) SPX_LP(2, __nv_bfloat16, true); else SPX_LP(2, __nv_bfloat16, false); }
        else { if (pre) SPX_LP(2, float, true); else SPX_LP(2, float, false); }
    } else if (R == 4) {
        if (dtype_bf16) { if (pre) SPX_LP(4, __nv_bfloat16, true); else SPX_LP(4, __nv_bfloat16, false); }
        else { if (pre) SPX_LP(4, float, true); else SPX_LP(4, float, false); }
    }
#undef SPX_LP
    return cudaErrorInvalidValue;
}

}  // namespace spx
