// spectre_gate.cu -- stand-alone gate expansion for ALL heads of a layer in one launch:
//   anchors (B, NG, Bk) complex64  ->  gate_half (B, NG, F_half) complex64
// replacing, per head, interp_complex_1d (spectre.py:526-528 -> :26-61), ComplexModReLU (:531 -> :109-121) and the
// positional phase (:534-536): about ten small PyTorch kernels per head and forward.
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <tuple>

#include "../../include/spectre_mix.h"
#include "spectre_gate.cuh"
#include "spectre_internal.h"

namespace {

// flat index over (b, g, k): no grid.y / grid.z limits, any B and NG the reference accepts
// DECODE: multiply by the positional phase of SpectreHead.decode_step (spectre.py:594-598), exp(j theta_k) with
// theta_k = fl(fl(fl(2 pi32 * k) * (t - j)) / n) -- complex64 tensor arithmetic of the reference, in its rounding order
template <bool DECODE>
__global__ void __launch_bounds__(256) gate_expand_kernel(spx::GateSrc s, float2 *__restrict__ gate, int NG, int F_half, long long total,
                                                          float t_minus_j, float n_fft) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i % F_half);
        const long long r = i / F_half;
        const int g = (int)(r % NG);
        const long long b = r / NG;
        const int head = g / s.G, j = g - head * s.G;
        const float2 *a = s.anchors + ((size_t)b * NG + (size_t)head * s.G) * s.Bk;
        const float2 *pos = s.pos ? s.pos + (size_t)b * s.pos_stride_b : nullptr;
        float2 z = spx::gate_from_anchors(a, s.Bk, s.G, j, k, F_half, __ldg(s.bias + (size_t)g * F_half + k), __ldg(s.eps + g), pos);
        if (DECODE) {
            const float two_pi32 = (float)(2.0 * M_PI);
            float sn, cs;
            sincosf(__fdiv_rn(__fmul_rn(__fmul_rn(two_pi32, (float)k), t_minus_j), n_fft), &sn, &cs);
            z = make_float2(z.x * cs - z.y * sn, z.x * sn + z.y * cs);
        }
        gate[i] = z;
    }
}

int check_args(const void *anchors, const float *bias, const float *eps, void *gate, int B, int NG, int G, int Bk, int F_half) {
    if (!anchors || !bias || !eps || !gate)
        return spx::fail(SPECTRE_MIX_ERR_BAD_ARG, "gate expand: null pointer (anchors=%p bias=%p eps=%p gate=%p)", anchors, (const void *)bias,
                         (const void *)eps, gate);
    if (G <= 0 || NG <= 0 || NG % G != 0) return spx::fail(SPECTRE_MIX_ERR_BAD_ARG, "gate expand: NG=%d is not a positive multiple of G=%d", NG, G);
    if (B < 0 || Bk < 1 || F_half < 2) return spx::fail(SPECTRE_MIX_ERR_BAD_ARG, "gate expand: bad size (B=%d Bk=%d F_half=%d)", B, Bk, F_half);
    return 0;
}

int grid_for(long long total) { return (int)std::min<long long>((total + 255) / 256, 1 << 20); }

__global__ void gate_interp_table_kernel(int F_half, int Bk, float4 *coef, ushort4 *tap) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= F_half) return;
    float4 c;
    int tp[4];
    spx::gate_interp_of(k, F_half, Bk, c, tp);
    coef[k] = c;
    tap[k] = make_ushort4((unsigned short)tp[0], (unsigned short)tp[1], (unsigned short)tp[2], (unsigned short)tp[3]);
}

struct InterpTable {
    float4 *coef = nullptr;
    ushort4 *tap = nullptr;
};
std::mutex g_tab_mu;
std::map<std::tuple<int, int, int>, InterpTable> g_tabs;   // (device, F_half, Bk)

}  // namespace

namespace spx {
int gate_interp_table(int F_half, int Bk, const float4 **icoef, const ushort4 **itap) {
    *icoef = nullptr;
    *itap = nullptr;
    if (Bk > 65535) return 0;   // taps do not fit 16 bits: the kernel derives them per bin instead
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    std::lock_guard<std::mutex> lock(g_tab_mu);
    auto key = std::make_tuple(dev, F_half, Bk);
    auto it = g_tabs.find(key);
    if (it == g_tabs.end()) {
        InterpTable t;
        if ((e = cudaMalloc(&t.coef, sizeof(float4) * (size_t)F_half)) != cudaSuccess) return cuda_fail(e, "cudaMalloc(interp table)");
        if ((e = cudaMalloc(&t.tap, sizeof(ushort4) * (size_t)F_half)) != cudaSuccess) { cudaFree(t.coef); return cuda_fail(e, "cudaMalloc(interp table)"); }
        gate_interp_table_kernel<<<(F_half + 255) / 256, 256>>>(F_half, Bk, t.coef, t.tap);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaDeviceSynchronize();   // first use only: the table is complete before anyone reads it
        if (e != cudaSuccess) { cudaFree(t.coef); cudaFree(t.tap); return cuda_fail(e, "interpolation table kernel"); }
        it = g_tabs.emplace(key, t).first;
    }
    *icoef = it->second.coef;
    *itap = it->second.tap;
    return 0;
}
}  // namespace spx

extern "C" int spectre_gate_expand(const void *anchors, const float *bias, const float *eps, const void *pos_phase,
                                   long long pos_stride_b, void *gate, int B, int NG, int G, int Bk, int F_half, void *stream) {
    if (int rc = check_args(anchors, bias, eps, gate, B, NG, G, Bk, F_half)) return rc;
    if (B == 0) return 0;
    spx::GateSrc s{reinterpret_cast<const float2 *>(anchors), bias, eps, reinterpret_cast<const float2 *>(pos_phase), pos_stride_b, Bk, G};
    const long long total = (long long)B * NG * F_half;
    gate_expand_kernel<false><<<grid_for(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(s, reinterpret_cast<float2 *>(gate), NG, F_half,
                                                                                                   total, 0.f, 1.f);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : spx::cuda_fail(e, "gate expand kernel launch");
}

extern "C" int spectre_decode_gate(const void *anchors, const float *bias, const float *eps, void *gate, int NG, int G, int Bk,
                                   int F_half, long long t, int n_fft, void *stream) {
    if (int rc = check_args(anchors, bias, eps, gate, 1, NG, G, Bk, F_half)) return rc;
    if (n_fft < 2 || t < 0) return spx::fail(SPECTRE_MIX_ERR_BAD_ARG, "decode gate: bad n_fft=%d / t=%lld", n_fft, t);
    spx::GateSrc s{reinterpret_cast<const float2 *>(anchors), bias, eps, nullptr, 0, Bk, G};
    const long long total = (long long)NG * F_half;
    const float tmj = (float)(t - t % n_fft);
    gate_expand_kernel<true><<<grid_for(total), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(s, reinterpret_cast<float2 *>(gate), NG, F_half,
                                                                                                  total, tmj, (float)n_fft);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : spx::cuda_fail(e, "decode gate kernel launch");
}
