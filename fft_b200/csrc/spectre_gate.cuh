// spectre_gate.cuh -- gate generator tail: anchors -> cubic interpolation -> modReLU (-> positional phase).
//
// Stands in for spectre.py:526-536:
//   interp_complex_1d(gate_anchor, size=F_half, mode="cubic")   spectre.py:26-61   (grid_sample, bicubic, border, align_corners)
//   ComplexModReLU                                              spectre.py:109-121 (relu(|z| + b) * z / sqrt(|z|^2 + eps^2))
//   gate_half * pos_phase                                       spectre.py:534-536
// One device function evaluates one frequency bin of one gate row, so that the stand-alone expand kernel and the mix
// kernel's gate staging share the arithmetic.  The interpolation follows ATen's grid sampler: sampling positions are
// torch.linspace(-1, 1, F_half) un-normalised with align_corners=True, taps floor-1 .. floor+2 clamped to the border,
// cubic-convolution coefficients with A = -0.75, float32 throughout.
//
// Reference quirk that parity requires (spectre.py:41): the anchors are stacked as (B, 2, G, K) and then RESHAPED to
// (B*G, 2, 1, K), so the "real" and "imaginary" planes handed to grid_sample are consecutive rows of the list
//   R = [re_0, ..., re_{G-1}, im_0, ..., im_{G-1}]      (per sample and head)
// i.e. gate row j of a head gets  real = interp(R[2j]),  imag = interp(R[2j+1]).  With G = 4: row 0 = (re_0, re_1),
// row 1 = (re_2, re_3), row 2 = (im_0, im_1), row 3 = (im_2, im_3).  The kernel reproduces exactly that.
#pragma once
#include <cuda_runtime.h>

namespace spx {

struct GateSrc {
    const float2 *anchors;   // [B][NG][Bk] complex64
    const float *bias;       // [NG][F_half] modReLU bias (heads stacked)
    const float *eps;        // [NG] modReLU epsilon per gate row
    const float2 *pos;       // nullable positional phase, complex64 [*][F_half]
    long long pos_stride_b;  // 0 = one phase row shared by the batch
    int Bk;                  // anchors per gate row
    int G;                   // gate rows per head (the scope of the reference's real/imag row shuffle)
    // Optional per-(F_half, Bk) interpolation table (built once per device by the library, spectre_gate.cu): the four cubic
    // coefficients and the four clamped tap indices of every bin depend on k only, so the mix kernel's gate staging reads
    // them (24 B per bin, L2-resident) instead of re-deriving them for every (batch row, gate row).  nullptr = compute.
    const float4 *icoef;     // [F_half] (c0, c1, c2, c3)
    const ushort4 *itap;     // [F_half] (t0, t1, t2, t3)
};

__device__ __forceinline__ float cubic_conv1(float x, float A) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; }
__device__ __forceinline__ float cubic_conv2(float x, float A) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; }

// one float plane of the shuffled anchor list: row r (< 2G) of head block `a` ([G][Bk] complex): component r / G of row r % G
__device__ __forceinline__ float plane_tap(const float2 *__restrict__ a, int Bk, int G, int r, int i) {
    const float *p = reinterpret_cast<const float *>(a + (size_t)(r % G) * Bk + i);
    return __ldg(p + r / G);
}

// cubic-convolution coefficients and clamped taps of bin k: ATen's grid sampler on torch.linspace(-1, 1, F_half)
__device__ __forceinline__ void gate_interp_of(int k, int F_half, int Bk, float4 &c, int (&tp)[4]) {
    // torch.linspace(-1, 1, F_half): symmetric evaluation around the middle
    const float step = 2.f / (float)(F_half - 1);
    const float gx = (k < F_half / 2) ? -1.f + step * (float)k : 1.f - step * (float)(F_half - 1 - k);
    const float ix = ((gx + 1.f) * 0.5f) * (float)(Bk - 1);
    const float fl = floorf(ix);
    const float t = ix - fl;
    const int i1 = (int)fl;
    constexpr float A = -0.75f;
    c = make_float4(cubic_conv2(t + 1.f, A), cubic_conv1(t, A), cubic_conv1(1.f - t, A), cubic_conv2(2.f - t, A));
    const int hi = Bk - 1;
    tp[0] = min(max(i1 - 1, 0), hi); tp[1] = min(max(i1, 0), hi); tp[2] = min(max(i1 + 1, 0), hi); tp[3] = min(max(i1 + 2, 0), hi);
}

// interpolated anchors of gate row j at bin k (before modReLU); coefficients in the grid sampler's summation order
__device__ __forceinline__ float2 gate_interp_eval(const float2 *__restrict__ a, int Bk, int G, int j, const float4 &c, const int (&tp)[4]) {
    const int r0 = 2 * j, r1 = 2 * j + 1;
    float2 z;
    z.x = plane_tap(a, Bk, G, r0, tp[0]) * c.x + plane_tap(a, Bk, G, r0, tp[1]) * c.y + plane_tap(a, Bk, G, r0, tp[2]) * c.z +
          plane_tap(a, Bk, G, r0, tp[3]) * c.w;
    z.y = plane_tap(a, Bk, G, r1, tp[0]) * c.x + plane_tap(a, Bk, G, r1, tp[1]) * c.y + plane_tap(a, Bk, G, r1, tp[2]) * c.z +
          plane_tap(a, Bk, G, r1, tp[3]) * c.w;
    return z;
}

// gate value of bin k (< F_half) of gate row j (< G) of one (sample, head); `a` = that head's [G][Bk] anchors; bias / eps /
// pos already offset to the row.  The stand-alone expansion kernel: IEEE square root and division as ATen evaluates them.
__device__ __forceinline__ float2 gate_from_anchors(const float2 *__restrict__ a, int Bk, int G, int j, int k, int F_half,
                                                    float bias, float eps, const float2 *__restrict__ pos) {
    float4 c;
    int tp[4];
    gate_interp_of(k, F_half, Bk, c, tp);
    float2 z = gate_interp_eval(a, Bk, G, j, c, tp);
    const float mag = sqrtf(z.x * z.x + z.y * z.y);
    const float scale = fmaxf(mag + bias, 0.f) / sqrtf(mag * mag + eps * eps);
    z.x *= scale;
    z.y *= scale;
    if (pos) {
        const float2 p = __ldg(pos + k);
        z = make_float2(z.x * p.x - z.y * p.y, z.x * p.y + z.y * p.x);
    }
    return z;
}

// The same function for the mix kernel's gate staging, where it runs once per (tile, bin) on the kernel's critical path: the
// k-only part comes from the table when the library built one, and modReLU uses the reciprocal square root unit
// (|z| = m2 * rsqrt(m2), scale = relu(|z| + b) * rsqrt(m2 + eps^2); MUFU.RSQ is accurate to 2 ulp -- well inside the 1e-5 bar,
// checked against the reference-generated gate goldens through the fused entry).
__device__ __forceinline__ float2 gate_from_anchors_fast(const float4 *__restrict__ icoef, const ushort4 *__restrict__ itap, int Bk, int G,
                                                         const float2 *__restrict__ a, int j, int k, int F_half, float bias, float eps,
                                                         const float2 *__restrict__ pos) {
    float4 c;
    int tp[4];
    if (icoef) {
        c = __ldg(icoef + k);
        const ushort4 t = __ldg(itap + k);
        tp[0] = t.x; tp[1] = t.y; tp[2] = t.z; tp[3] = t.w;
    } else {
        gate_interp_of(k, F_half, Bk, c, tp);
    }
    float2 z = gate_interp_eval(a, Bk, G, j, c, tp);
    const float m2 = z.x * z.x + z.y * z.y;
    const float mag = m2 * rsqrtf(fmaxf(m2, 1e-37f));
    const float scale = fmaxf(mag + bias, 0.f) * rsqrtf(fmaf(eps, eps, m2));
    z.x *= scale;
    z.y *= scale;
    if (pos) {
        const float2 p = __ldg(pos + k);
        z = make_float2(z.x * p.x - z.y * p.y, z.x * p.y + z.y * p.x);
    }
    return z;
}

}  // namespace spx
