// spectre_gate.cu -- stand-alone gate expansion for ALL heads of a layer in one launch:
//   anchors (B, NG, Bk) complex64  ->  gate_half (B, NG, F_half) complex64
// replacing, per head, interp_complex_1d (spectre.py:526-528 -> :26-61), ComplexModReLU (:531 -> :109-121) and the
// positional phase (:534-536): about ten small PyTorch kernels per head and forward.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/spectre_mix.h"
#include "spectre_gate.cuh"

namespace {

__global__ void __launch_bounds__(256) gate_expand_kernel(spx::GateSrc s, float2 *__restrict__ gate, int NG, int F_half) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int g = blockIdx.y, b = blockIdx.z;
    if (k >= F_half) return;
    const int head = g / s.G, j = g - head * s.G;
    const float2 *a = s.anchors + ((size_t)b * NG + (size_t)head * s.G) * s.Bk;
    const float2 *pos = s.pos ? s.pos + (size_t)b * s.pos_stride_b : nullptr;
    gate[((size_t)b * NG + g) * F_half + k] =
        spx::gate_from_anchors(a, s.Bk, s.G, j, k, F_half, __ldg(s.bias + (size_t)g * F_half + k), __ldg(s.eps + g), pos);
}

}  // namespace

extern "C" int spectre_gate_expand(const void *anchors, const float *bias, const float *eps, const void *pos_phase,
                                   long long pos_stride_b, void *gate, int B, int NG, int G, int Bk, int F_half, void *stream) {
    if (!anchors || !bias || !eps || !gate) return SPECTRE_MIX_ERR_BAD_ARG;
    if (G <= 0 || NG % G != 0) return SPECTRE_MIX_ERR_BAD_ARG;
    if (B < 0 || NG <= 0 || Bk < 1 || F_half < 2 || NG > 65535 || B > 65535) return SPECTRE_MIX_ERR_BAD_ARG;
    if (B == 0) return 0;
    spx::GateSrc s{reinterpret_cast<const float2 *>(anchors), bias, eps, reinterpret_cast<const float2 *>(pos_phase), pos_stride_b, Bk, G};
    dim3 grid((F_half + 255) / 256, NG, B);
    gate_expand_kernel<<<grid, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(s, reinterpret_cast<float2 *>(gate), NG, F_half);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? 0 : SPECTRE_MIX_ERR_CUDA + (int)e;
}
