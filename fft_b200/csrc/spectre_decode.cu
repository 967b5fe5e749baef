// spectre_decode.cu -- autoregressive decode side of the Spectre mixer.
//
// Replaces, for ALL heads of a layer at once (channels = embed_dim, gate row of channel c = c / group_width):
//   * PrefixFFTCache.decode_step   spectre.py:795-806  running spectrum += e^{j w k t} v_t  (- e^{j w k j} v_old when evicting)
//   * SpectreHead.decode_step      spectre.py:605      mixed_half = gate_broadcast * prefix_fft
//   * pruned_irfft_single          spectre.py:614-655  one output sample of the inverse real FFT
// as ONE bandwidth-bound pass over prefix_fft (F_half x d complex64): read, update, write, and reduce over k.
//
// The reference evaluates every phase angle in float32 (complex64 tensor arithmetic); the kernels reproduce that
// rounding order so that results agree to float32 round-off rather than to the exact angles:
//   cache phases   theta = fl(fl(w32 * k) * t),            w32 = fl32(-2 pi / n_fft)        (spectre.py:766, :801, :805)
//   readout phase  phi   = fl(fl(fl(2pi32 * k) * pos) / n)                                   (spectre.py:628)
// and keep the reference's Nyquist term contrib[-1] * (-1)^pos (spectre.py:650) as it is.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/spectre_mix.h"
#include "spectre_internal.h"

namespace {

#ifndef SPX_DECODE_KCHUNK
#define SPX_DECODE_KCHUNK 16
#endif
constexpr int kKChunk = SPX_DECODE_KCHUNK;  // frequency bins per block
#ifndef SPX_DECODE_UNROLL
#define SPX_DECODE_UNROLL 8
#endif
constexpr int kUnroll = SPX_DECODE_UNROLL;   // bins in flight per thread

// sin / cos of a float32 angle of any size the decode path produces (|theta| up to ~1e5 rad): two-constant Cody-Waite
// reduction by 2 pi with fused multiply-adds (exact to ~1e-7 rad for |n| < 2^16), then the special-function unit on
// [-pi, pi] (absolute error ~5e-7).  The ANGLE keeps the reference's float32 rounding order (that is what parity needs: at
// 1e4 rad a float32 angle is only known to 1e-3 rad, and the reference uses exactly that value); its sine and cosine need
// not be correctly rounded -- sincosf costs ~4x the instructions on the kernel's critical path.
__device__ __forceinline__ void sincos_reduced(float theta, float *s, float *c) {
    const float n = rintf(theta * 0.15915494309189535f);
    float r = fmaf(n, -6.2831854820251465f, theta);      // 2 pi = 6.2831854820251465 - 1.7484555e-7
    r = fmaf(n, 1.7484555e-7f, r);
    *s = __sinf(r);
    *c = __cosf(r);
}

// mode bit 0: update the spectrum; bit 1: read one output sample out
template <int MODE>
__global__ void __launch_bounds__(256) decode_kernel(float2 *__restrict__ prefix, const float *__restrict__ v_new,
                                                     const float *__restrict__ v_old, const float2 *__restrict__ gate,
                                                     float *__restrict__ partial, int n_fft, int d, int group_width, float t_new,
                                                     float t_old, int evict, int pos, float w32) {
    // programmatic dependent launch: the blocks may become resident while the previous kernel of the stream (the decode gate, the
    // previous token's reduction) still runs; nothing is read or written before the wait, and the next launch may follow likewise
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int c = blockIdx.y * blockDim.x + threadIdx.x;
    if (c >= d) return;
    const int F_half = n_fft / 2 + 1;
    const int k0 = blockIdx.x * kKChunk;
    const int k1 = min(k0 + kKChunk, F_half);
    float vn = 0.f, vo = 0.f;
    if (MODE & 1) {
        vn = v_new[c];
        if (evict) vo = v_old[c];
    }
    const float2 *grow = (MODE & 2) ? gate + (size_t)(c / group_width) * F_half : nullptr;
    const float two_pi32 = (float)(2.0 * M_PI);
    const float posf = (float)pos, nf = (float)n_fft;
    const float nyq_sign = (pos & 1) ? -1.f : 1.f;
    float acc = 0.f;
#pragma unroll kUnroll
    for (int k = k0; k < k1; ++k) {
        float2 X = prefix[(size_t)k * d + c];
        const float kf = (float)k;
        if (MODE & 1) {
            const float a = __fmul_rn(w32, kf);
            if (evict) {   // spectre.py:799-802: subtract the evicted token first
                float s, co;
                sincos_reduced(__fmul_rn(a, t_old), &s, &co);
                X.x -= co * vo;
                X.y -= s * vo;
            }
            float s, co;
            sincos_reduced(__fmul_rn(a, t_new), &s, &co);
            X.x += co * vn;
            X.y += s * vn;
            prefix[(size_t)k * d + c] = X;
        }
        if (MODE & 2) {
            const float2 g = grow[k];
            const float yr = g.x * X.x - g.y * X.y;           // gate_broadcast * prefix_fft
            const float yi = g.x * X.y + g.y * X.x;
            float s, co;
            sincos_reduced(__fdiv_rn(__fmul_rn(__fmul_rn(two_pi32, kf), posf), nf), &s, &co);
            const float contrib = yr * co - yi * s;
            float wgt = 2.f;                                    // spectre.py:643-653
            if (k == 0) wgt = 1.f;
            else if (k == F_half - 1 && (n_fft % 2) == 0) wgt = nyq_sign;
            acc += wgt * contrib;
        }
    }
    // deterministic read-out: every block leaves its partial sum, decode_reduce_kernel adds them in a fixed order
    if (MODE & 2) partial[(size_t)blockIdx.x * d + c] = acc;
}

// out[c] = (sum over frequency chunks) / n_fft in a FIXED order, so a token is bit-reproducible run to run: a block serves 32
// channels with 8 slices of threads; slice s adds the partials of chunks s, s + 8, s + 16, ... in order (16 dependent loads per
// thread at n_fft = 4096 instead of 129), then thread (0, c) adds the 8 slice sums in slice order.
constexpr int kRedCh = 32, kRedSlices = 8;
__global__ void __launch_bounds__(kRedCh * kRedSlices) decode_reduce_kernel(const float *__restrict__ partial, float *__restrict__ out,
                                                                            int nchunks, int d, float nf) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    __shared__ float part_s[kRedSlices][kRedCh];
    const int cl = threadIdx.x % kRedCh, sl = threadIdx.x / kRedCh;
    const int c = blockIdx.x * kRedCh + cl;
    float acc = 0.f;
    if (c < d) {
#pragma unroll 4
        for (int i = sl; i < nchunks; i += kRedSlices) acc += partial[(size_t)i * d + c];
    }
    part_s[sl][cl] = acc;
    __syncthreads();
    if (sl == 0 && c < d) {
        float t = part_s[0][cl];
#pragma unroll
        for (int j = 1; j < kRedSlices; ++j) t += part_s[j][cl];
        out[c] = t / nf;
    }
}

int nchunks_of(int n_fft) { return (n_fft / 2 + 1 + kKChunk - 1) / kKChunk; }

int launch(int mode, void *prefix, const float *v_new, const float *v_old, const void *gate, float *out, int n_fft, int d,
           int group_width, long long t, int pos, void *ws, size_t ws_bytes, cudaStream_t st) {
    if (n_fft < 2 || d <= 0 || group_width <= 0 || (d % group_width) != 0)
        return spx::fail(SPECTRE_MIX_ERR_BAD_ARG, "decode: bad size (n_fft=%d d=%d group_width=%d)", n_fft, d, group_width);
    const int F_half = n_fft / 2 + 1;
    const int threads = d >= 256 ? 256 : ((d + 31) / 32) * 32;
    const int nchunks = nchunks_of(n_fft);
    dim3 grid(nchunks, (d + threads - 1) / threads);
    (void)F_half;
    const float w32 = (float)(-2.0 * M_PI / (double)n_fft);
    const long long j = t % n_fft;
    const int evict = (t >= n_fft) ? 1 : 0;
    if (mode & 2) {
        const size_t need = (size_t)nchunks * d * sizeof(float);
        if (!ws || ws_bytes < need)
            return spx::fail(SPECTRE_MIX_ERR_BAD_ARG, "decode: workspace %zu bytes given, spectre_decode_workspace_bytes() = %zu", ws ? ws_bytes : 0, need);
    }
    float2 *pf = reinterpret_cast<float2 *>(prefix);
    const float2 *g = reinterpret_cast<const float2 *>(gate);
    float *part = reinterpret_cast<float *>(ws);
    // both launches carry the programmatic-stream-serialization attribute (the kernels wait with griddepcontrol.wait before
    // they touch memory): a decode step is two dependent 25 MB / 0.4 MB passes, i.e. launch-latency bound
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(threads);
    cfg.stream = st;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const float tf = (float)t, jf = (float)j;
    cudaError_t e;
    switch (mode) {
        case 1: e = cudaLaunchKernelEx(&cfg, decode_kernel<1>, pf, v_new, v_old, g, part, n_fft, d, group_width, tf, jf, evict, pos, w32); break;
        case 2: e = cudaLaunchKernelEx(&cfg, decode_kernel<2>, pf, v_new, v_old, g, part, n_fft, d, group_width, tf, jf, evict, pos, w32); break;
        default: e = cudaLaunchKernelEx(&cfg, decode_kernel<3>, pf, v_new, v_old, g, part, n_fft, d, group_width, tf, jf, evict, pos, w32); break;
    }
    if (e != cudaSuccess) return spx::cuda_fail(e, "decode kernel launch");
    if (mode & 2) {
        cfg.gridDim = dim3((d + kRedCh - 1) / kRedCh);
        cfg.blockDim = dim3(kRedCh * kRedSlices);
        e = cudaLaunchKernelEx(&cfg, decode_reduce_kernel, (const float *)part, out, nchunks, d, (float)n_fft);
        if (e != cudaSuccess) return spx::cuda_fail(e, "decode reduce kernel launch");
    }
    return 0;
}

}  // namespace

extern "C" {

size_t spectre_decode_workspace_bytes(int n_fft, int d) {
    if (n_fft < 2 || d <= 0) return 0;
    return (size_t)nchunks_of(n_fft) * (size_t)d * sizeof(float);
}

int spectre_decode_update(void *prefix_fft, const float *v_new, const float *v_old, int n_fft, int d, long long t, void *stream) {
    if (!prefix_fft || !v_new || (t >= n_fft && !v_old)) return spx::fail(SPECTRE_MIX_ERR_BAD_ARG, "decode update: null pointer");
    return launch(1, prefix_fft, v_new, v_old, nullptr, nullptr, n_fft, d, 1, t, 0, nullptr, 0, reinterpret_cast<cudaStream_t>(stream));
}

int spectre_decode_readout(const void *prefix_fft, const void *gate, float *out, int n_fft, int d, int group_width, int pos,
                           void *workspace, size_t workspace_bytes, void *stream) {
    if (!prefix_fft || !gate || !out) return spx::fail(SPECTRE_MIX_ERR_BAD_ARG, "decode readout: null pointer");
    return launch(2, const_cast<void *>(prefix_fft), nullptr, nullptr, gate, out, n_fft, d, group_width, 0, pos, workspace,
                  workspace_bytes, reinterpret_cast<cudaStream_t>(stream));
}

int spectre_decode_step(void *prefix_fft, const float *v_new, const float *v_old, const void *gate, float *out, int n_fft,
                        int d, int group_width, long long t, void *workspace, size_t workspace_bytes, void *stream) {
    if (!prefix_fft || !v_new || !gate || !out || (t >= n_fft && !v_old)) return spx::fail(SPECTRE_MIX_ERR_BAD_ARG, "decode step: null pointer");
    return launch(3, prefix_fft, v_new, v_old, gate, out, n_fft, d, group_width, t, (int)(t % n_fft), workspace, workspace_bytes,
                  reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
