// spectre_mix_kernel.cuh -- fused rFFT -> gate (+memory) -> irFFT for sm_100a.
//
// Replaces /root/reference/spectre.py:506 (torch.fft.rfft), :542-545 (gate
// broadcast + complex multiply), :548-549 (memory add), :551 (torch.fft.irfft),
// :553 ([:N]) and the per-head loop of :703-718 with ONE persistent kernel.
//
// How the path is mapped onto B200 (details in DESIGN.md):
//   * The transform axis is the strided one (V is [B][N][C], channels contiguous),
//     so a CTA owns ALL n_fft rows of a narrow channel tile of one batch row and
//     keeps them in shared memory for the whole rfft->gate->irfft round trip:
//     HBM sees each input element once and each output element once.
//   * Four adjacent channels (c..c+3, same gate group) form one tile "element":
//     two complex signals z0 = v[c] + i v[c+2], z1 = v[c+1] + i v[c+3].  Because the
//     gate acts as a real convolution kernel (irfft drops imag(DC), imag(Nyquist)),
//     ifft(Gfull * fft(z)) returns the two real results in Re and Im -- half the
//     FFT work and no real-FFT post-processing.  z0 and z1 ride in the two lanes
//     of Blackwell's packed fp32x2 pipe (FADD2 / FMUL2 / FFMA2): every butterfly
//     instruction does two columns, twiddles are broadcast scalar operands.
//   * FFT = in-place mixed radix (r,16,16[,16]) decimation in frequency going
//     forward, the exact mirror (decimation in time) coming back, so no digit
//     reversal pass exists: the gate is applied in digit-reversed order, and the
//     last forward butterfly, the gate multiply and the first inverse butterfly
//     are one register-resident pass.
//   * Inverse transform = forward butterflies on (im, re)-swapped data, so one
//     twiddle table serves both directions.
//   * Shared-memory element index e is stored at e + (e >> 4): every pass (strides
//     n_fft/r, 16, 1) is bank-conflict free with 128-bit accesses.
//   * Tile I/O goes through TMA (cp.async.bulk.tensor) wherever the layout allows; at n_fft = 4096 a helper
//     warpgroup parks the NEXT tile in tensor memory and drains the PREVIOUS tile's results from it through one
//     ring of TMA boxes while the 16 compute warps run the FFT passes (TMEM_IO variant, fp32 and bf16 rows).
//   * A radix-16 stage keeps 6 of its 15 twiddle rows in shared memory and forms the rest as scalar products.
//   * Compute warps are staggered per scheduler slot after the CTA barriers and the barrier around the last
//     pass's read is split (arrive / wait), so shared-memory phases of one warp run under another's butterflies.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#include "spectre_gate.cuh"

namespace spx {

enum Mode : int { MODE_QUAD = 0, MODE_PAIR = 1, MODE_REAL = 2 };

struct MixParams {
    const void *v;
    const float2 *gate;   // [B][NG][F_half]
    const float2 *mem;    // [F_half][C] (row stride mem_stride) or nullptr
    void *out;            // mix: [B][n_out][C] real; rfft-only: complex64 [B][F_half][C]
    const float2 *tw;     // per-plan twiddle table (device)
    long long v_sb, v_sn, o_sb, o_sn, mem_stride;
    int B, n_in, n_out, C, group_width, NG;
    int tiles_per_row, num_tiles;
    int gate_tables;      // gate groups a tile may touch (smem sized for this many)
    float inv_n;
    int prefetch;         // 1: prefetch the CTA's next tile into L2 while this one is transformed
    int gw_shift;         // log2(group_width) when it is a power of two, else -1 (the kernel then divides)
    int sub_shift;        // log2(sub_R)
    int sub_R;            // long-context path: number of interleaved sub-transforms (n_total = sub_R * n_fft), else 1
    int skew_ns;          // warp stagger code (see stagger()): 0 off, > 0 legacy nanosleep, < 0 clock spin per scheduler slot
    int sched;            // bit 0: stagger also after the barrier before inverse stage 0; bit 1 (TMEM variant): split barrier
                          // around the inverse stage-0 read (arrive after the read, wait before the next tile's stage-0 write)
    unsigned long long *timeline;  // optional: per-CTA phase timestamps (ns) for tools/timeline.py, else nullptr
    int pair_tiles;       // TMEM variant: a CTA takes the two channel tiles of one gate group back to back (tiles 2P, 2P+1 of pair
                          // P = blockIdx.x + k gridDim.x) and stages the gate row once for both; the host sets it only when
                          // tiles_per_row is even and one gate row covers both tiles
    // ANCH kernels (spectre_mix_fwd_anchors): the gate is never materialised -- the gate staging evaluates
    // cubic interpolation + modReLU (+ positional phase) of spectre.py:526-536 from the anchors of the row's head
    GateSrc gsrc;                  // anchors [B][NG][Bk], bias [NG][F_half], eps [NG], optional phase; gate == nullptr
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
constexpr int kTimelineSlots = 8;     // timestamps per tile
constexpr int kTimelineGroups = 5;    // four compute thread groups + the helper warpgroup's elected thread
constexpr int kTimelineTiles = 8;     // tiles recorded per CTA (by thread 0 of each of the up to four 128-thread groups)

// ------------------------------------------------------------------ packed / scalar lanes
__device__ __forceinline__ float2 vadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 vsub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 vmul(float2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
__device__ __forceinline__ float2 vfma(float2 a, float s, float2 c) { return __ffma2_rn(a, make_float2(s, s), c); }
__device__ __forceinline__ float2 vfnma(float2 a, float s, float2 c) { return __ffma2_rn(a, make_float2(-s, -s), c); }
__device__ __forceinline__ float2 vzero(float2) { return make_float2(0.f, 0.f); }
__device__ __forceinline__ float2 vneg(float2 a) { return make_float2(-a.x, -a.y); }

__device__ __forceinline__ float vadd(float a, float b) { return a + b; }
__device__ __forceinline__ float vsub(float a, float b) { return a - b; }
__device__ __forceinline__ float vmul(float a, float s) { return a * s; }
__device__ __forceinline__ float vfma(float a, float s, float c) { return fmaf(a, s, c); }
__device__ __forceinline__ float vfnma(float a, float s, float c) { return fmaf(-a, s, c); }
__device__ __forceinline__ float vzero(float) { return 0.f; }
__device__ __forceinline__ float vneg(float a) { return -a; }

template <class V>
struct Cx {
    V re, im;
};

template <class V>
__device__ __forceinline__ Cx<V> cadd(const Cx<V> &a, const Cx<V> &b) { return {vadd(a.re, b.re), vadd(a.im, b.im)}; }
template <class V>
__device__ __forceinline__ Cx<V> csub(const Cx<V> &a, const Cx<V> &b) { return {vsub(a.re, b.re), vsub(a.im, b.im)}; }
// a * (wr + i wi), wr/wi scalars shared by both packed lanes
template <class V>
__device__ __forceinline__ Cx<V> cmul(const Cx<V> &a, float wr, float wi) {
    Cx<V> r;
    r.re = vfnma(a.im, wi, vmul(a.re, wr));
    r.im = vfma(a.im, wr, vmul(a.re, wi));
    return r;
}
// a * conj(wr + i wi)
template <class V>
__device__ __forceinline__ Cx<V> cmulc(const Cx<V> &a, float wr, float wi) {
    Cx<V> r;
    r.re = vfma(a.im, wi, vmul(a.re, wr));
    r.im = vfnma(a.re, wi, vmul(a.im, wr));
    return r;
}
template <class V>
__device__ __forceinline__ Cx<V> cswap(const Cx<V> &a) { return {a.im, a.re}; }

// ------------------------------------------------------------------ small DFTs (forward sign, natural order out)
template <class V>
__device__ __forceinline__ void dft2(Cx<V> &a, Cx<V> &b) {
    Cx<V> t = a;
    a = cadd(t, b);
    b = csub(t, b);
}
template <class V>
__device__ __forceinline__ void dft4(Cx<V> &x0, Cx<V> &x1, Cx<V> &x2, Cx<V> &x3) {
    Cx<V> t0 = cadd(x0, x2), t1 = csub(x0, x2), t2 = cadd(x1, x3), t3 = csub(x1, x3);
    x0 = cadd(t0, t2);
    x2 = csub(t0, t2);
    x1.re = vadd(t1.re, t3.im);  // t1 - i t3
    x1.im = vsub(t1.im, t3.re);
    x3.re = vsub(t1.re, t3.im);  // t1 + i t3
    x3.im = vadd(t1.im, t3.re);
}

template <int R, class V>
struct Dft;
template <class V>
struct Dft<2, V> {
    static __device__ __forceinline__ void run(Cx<V> (&x)[2]) { dft2(x[0], x[1]); }
};
template <class V>
struct Dft<4, V> {
    static __device__ __forceinline__ void run(Cx<V> (&x)[4]) { dft4(x[0], x[1], x[2], x[3]); }
};
template <class V>
struct Dft<8, V> {
    static __device__ __forceinline__ void run(Cx<V> (&x)[8]) {
        constexpr float h = 0.70710678118654752440f;
        dft4(x[0], x[2], x[4], x[6]);
        dft4(x[1], x[3], x[5], x[7]);
        x[3] = cmul(x[3], h, -h);                       // W8^1
        { Cx<V> t = x[5]; x[5].re = t.im; x[5].im = vneg(t.re); }  // W8^2 = -i
        x[7] = cmul(x[7], -h, -h);                      // W8^3
        Cx<V> y[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            dft2(x[2 * q], x[2 * q + 1]);
            y[q] = x[2 * q];
            y[q + 4] = x[2 * q + 1];
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) x[q] = y[q];
    }
};
template <class V>
struct Dft<16, V> {
    static __device__ __forceinline__ void run(Cx<V> (&x)[16]) {
        // n = n0 + 4 n1, q = q1 + 4 q0:  W16^{nq} = W4^{n1 q1} W16^{n0 q1} W4^{n0 q0}
        constexpr float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;
        constexpr float h = 0.70710678118654752440f;
#pragma unroll
        for (int n0 = 0; n0 < 4; ++n0) dft4(x[n0], x[n0 + 4], x[n0 + 8], x[n0 + 12]);
        // x[n0 + 4 q1] *= W16^{n0 q1}
        x[1 + 4] = cmul(x[1 + 4], c1, -s1);     // W^1
        x[1 + 8] = cmul(x[1 + 8], h, -h);       // W^2
        x[1 + 12] = cmul(x[1 + 12], s1, -c1);   // W^3
        x[2 + 4] = cmul(x[2 + 4], h, -h);       // W^2
        { Cx<V> t = x[2 + 8]; x[2 + 8].re = t.im; x[2 + 8].im = vneg(t.re); }  // W^4 = -i
        x[2 + 12] = cmul(x[2 + 12], -h, -h);    // W^6
        x[3 + 4] = cmul(x[3 + 4], s1, -c1);     // W^3
        x[3 + 8] = cmul(x[3 + 8], -h, -h);      // W^6
        x[3 + 12] = cmul(x[3 + 12], -c1, s1);   // W^9
        Cx<V> y[16];
#pragma unroll
        for (int q1 = 0; q1 < 4; ++q1) {
            dft4(x[4 * q1], x[4 * q1 + 1], x[4 * q1 + 2], x[4 * q1 + 3]);
#pragma unroll
            for (int q0 = 0; q0 < 4; ++q0) y[q1 + 4 * q0] = x[4 * q1 + q0];
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) x[q] = y[q];
    }
};

// ------------------------------------------------------------------ plan (compile time)
// SUBP: this transform is one of R interleaved sub-transforms of a longer one (n_total = R * N, first radix-R
// stage and its twiddles done by the pre/post pass kernels of the long-context path): the input is already complex,
// the gate is gathered with stride R and is not Hermitian within the sub-spectrum.
// Twiddle rows kept per stage.  A radix-16 stage stores only q in {1, 2, 3, 4, 8, 12} (rows 0..5): W^{u (q1 + 4 q0)} =
// W^{u q1} W^{4 u q0}, so six loads and nine scalar complex products replace fifteen loads -- 40 % of the table, which is
// what lets the stage-0 table of the 4096-point plan shrink from 30 KB to 12 KB of shared memory.
// SPX_TW4 (experiment): a radix-16 stage with a LONG table (L >= 256: stage 0 of the 4096-point plans) stores only q in {1, 2, 4, 8}
// and forms W^{3u} = W^{u} W^{2u}, W^{12u} = W^{4u} W^{8u} as two more scalar products: 4 KB less shared memory at 4096, which
// together with the 5 KB that were free is one more ring slot (8 instead of 7) for the TMEM-staged kernel.
#ifndef SPX_TW4
#define SPX_TW4 0
#endif
__host__ __device__ constexpr bool tw_rows4(int R, int L) { return SPX_TW4 != 0 && R == 16 && L >= 256; }
__host__ __device__ constexpr int tw_rows(int R, int L) { return R == 16 ? (tw_rows4(R, L) ? 4 : 6) : R - 1; }

// SUBP = 2 (DIT2): the tile's two element columns hold the EVEN and the ODD rows of a transform of length 2 N (same four
// channels); the two N-point sub-spectra are combined by one decimation-in-time radix-2 butterfly in the middle pass, right
// where the gate is applied, and split again before the inverse passes -- a 2 N-point transform at the shared-memory traffic
// of an N-point one (n_fft = 8192 in ONE pass over HBM).
template <int R0, int R1, int R2, int R3, int SUBP = 0>
struct Plan {
    static constexpr bool kSub = (SUBP == 1);
    static constexpr bool kDit = (SUBP == 2);
    static constexpr int N = R0 * R1 * R2 * R3;
    static constexpr int NS = (R3 > 1) ? 4 : ((R2 > 1) ? 3 : ((R1 > 1) ? 2 : 1));
    static_assert(NS >= 2, "at least two stages");
    __host__ __device__ static constexpr int R(int s) { return s == 0 ? R0 : (s == 1 ? R1 : (s == 2 ? R2 : R3)); }
    __host__ __device__ static constexpr int P(int s) { return s == 0 ? 1 : (s == 1 ? R0 : (s == 2 ? R0 * R1 : R0 * R1 * R2)); }
    __host__ __device__ static constexpr int L(int s) { return N / (P(s) * R(s)); }
    // twiddles of stage s (s < NS-1): W_{N/P}^{u q}, row j of the stage's table at TWOFF(s) + j L + u (see tw_rows)
    __host__ __device__ static constexpr int TWOFF(int s) { return s == 0 ? 0 : TWOFF(s - 1) + tw_rows(R(s - 1), L(s - 1)) * L(s - 1); }
    static constexpr int TWN = TWOFF(NS - 1);
    static constexpr int NPAD = N + (N >> 4);
    static constexpr int GPAD = (N / 2) + ((N / 2) >> 4) + 1;  // padded gate table length (float2)
};

__host__ __device__ constexpr int cmax(int a, int b) { return a > b ? a : b; }
// ring of TMA boxes shared by both directions of the TMEM variant: while a tile is parked all but one slot hold
// loads in flight (latency cover), while results are drained all of them are store sources
#ifndef SPX_SUB_TW_SMEM
#define SPX_SUB_TW_SMEM 1   // sub-transform variant: stage-0 twiddle rows in shared memory (12 KB) at the price of one ring slot
#endif
constexpr int kTmemSlotsMax = SPX_TW4 ? 8 : 7;
constexpr int kTmaBoxRows = 256;
// TMEM-staged variants: one ring slot = one TMA box of 512 tile elements (8 KB of fp32 rows) but at most 256 rows
__host__ __device__ constexpr int tmem_box_rows(int ncol) { return 512 / ncol < kTmaBoxRows ? 512 / ncol : kTmaBoxRows; }
// ring slots of a plan: the sub-transform variant holds a full-length gate table (two half-length slots) and has room for six,
// or for five next to its stage-0 twiddle rows (measured faster: profiles/r01d_ab_sub_twiddles.txt)
template <class PL>
__host__ __device__ constexpr int tmem_slots() { return (PL::kSub || PL::kDit) ? (SPX_SUB_TW_SMEM ? 5 : 6) : kTmemSlotsMax; }
// register split of the TMEM variant (512 compute + 128 helper threads, 96 per thread at launch = 61440 in the CTA pool)
#ifndef SPX_ILV
#define SPX_ILV 1
#endif
#ifndef SPX_SPLIT_DUTY
#define SPX_SPLIT_DUTY 1
#endif
#ifndef SPX_HELPER_TL
#define SPX_HELPER_TL 0
#endif
#ifndef SPX_COOP_PF
#define SPX_COOP_PF 0
#endif
#ifndef SPX_LAZY_STORE_WAIT
#define SPX_LAZY_STORE_WAIT 0   // helper: await the previous TMA store's read just before the NEXT step's barrier instead of right after the commit (measured: no change, profiles/r03h_ab_lazy_store_wait.txt)
#endif
#ifndef SPX_HELPER_GATE
#define SPX_HELPER_GATE 0   // TMEM kernels with a materialised gate and one table per tile: the helper warpgroup stages the gate rows.
                            // Correct, but the helper has ~6 % slack per tile: 10.8 us per tile against 9.2 (profiles/r03a_ab_helper_gate.txt)
#endif
#ifndef SPX_DIT_ASYNC
#define SPX_DIT_ASYNC 1   // DIT2 kernels with a materialised gate: rows go into the table by LDGSTS, unscaled (1/n applied in the middle pass)
#endif
#ifndef SPX_GATE_ASYNC
#define SPX_GATE_ASYNC 0   // next tile's gate row by LDGSTS straight into the table: 0 never, 1 wherever possible, 2 kernels with both tensor-memory
                           // exchanges, 3 TMEM kernels with 8- / 16-channel tiles and the row left RAW (1/n applied in the middle pass)
#endif
#ifndef SPX_TMEMX
#define SPX_TMEMX 0   // 4096-class TMEM kernels, exchanges through tensor memory: bit 0 stage 1 -> middle pass, bit 1 middle pass -> inverse stage 1.
                      // Correct and 31 % less shared-memory traffic, but no faster (the kernel is not bound there): see DESIGN 3.9.
                      // (Do not combine with the diagnostic launch flag sched bit 2: the helper would wait for an exchange that mode skips.)
#endif
#ifndef SPX_TMEM_COMPUTE_REGS
#define SPX_TMEM_COMPUTE_REGS 112
#endif
constexpr int kTmemComputeRegs = SPX_TMEM_COMPUTE_REGS, kTmemHelperRegs = (96 * 640 - SPX_TMEM_COMPUTE_REGS * 512) / 128;
static_assert(32 + 8 * kTmemSlotsMax <= 96, "ring barriers overlap the TMEM base slot");
static_assert(kTmemHelperRegs % 8 == 0 && kTmemHelperRegs >= 24, "setmaxnreg takes multiples of 8");

// Which stages keep their twiddles in shared memory: all of them while the tables fit beside the tile; from
// n_fft = 8192 the (largest) stage-0 table is read through L2 instead, at 16384 every table is.
template <class PL>
struct TwPolicy {
    static constexpr int FROM = (PL::N >= 16384) ? (PL::NS - 1) : ((PL::N >= 8192 || (PL::kSub && !SPX_SUB_TW_SMEM)) ? 1 : 0);
    static constexpr int SMEM_N = PL::TWN - PL::TWOFF(FROM);   // entries held in shared memory
};
template <class PL, int S_>
__device__ __forceinline__ float2 tw_get(const float2 *tw_s, const float2 *tw_g, int idx) {
    if constexpr (S_ >= TwPolicy<PL>::FROM) return tw_s[PL::TWOFF(S_) - PL::TWOFF(TwPolicy<PL>::FROM) + idx];
    else return __ldg(tw_g + PL::TWOFF(S_) + idx);
}

// radix-16 twiddles from the six stored rows: a[q1] = W^{u q1}, b[q0] = W^{4 u q0} (index 0 unused)
template <class V>
__device__ __forceinline__ void apply_twiddles16(Cx<V> (&x)[16], const float2 (&a)[4], const float2 (&b)[4]) {
#pragma unroll
    for (int q0 = 0; q0 < 4; ++q0)
#pragma unroll
        for (int q1 = 0; q1 < 4; ++q1) {
            const int q = q1 + 4 * q0;
            if (q == 0) continue;
            float2 w;
            if (q0 == 0) w = a[q1];
            else if (q1 == 0) w = b[q0];
            else w = make_float2(fmaf(-a[q1].y, b[q0].y, a[q1].x * b[q0].x), fmaf(a[q1].x, b[q0].y, a[q1].y * b[q0].x));
            x[q] = cmul(x[q], w.x, w.y);
        }
}

// x[q] *= W^{u q}, q = 1 .. R-1, for stage S_ (forward sign; the inverse passes call it on (im, re)-swapped data)
template <class PL, int S_, class V>
__device__ __forceinline__ void apply_twiddles(Cx<V> (&x)[PL::R(S_)], const float2 *tw_s, const float2 *tw_g, int u) {
    constexpr int R = PL::R(S_), L = PL::L(S_);
    if constexpr (R == 16) {
        float2 a[4], b[4];   // a[q1] = W^{u q1}, b[q0] = W^{4 u q0}
        if constexpr (tw_rows4(R, L)) {   // rows q = 1, 2, 4, 8; W^{3u} and W^{12u} are products
            a[1] = tw_get<PL, S_>(tw_s, tw_g, u);
            a[2] = tw_get<PL, S_>(tw_s, tw_g, L + u);
            b[1] = tw_get<PL, S_>(tw_s, tw_g, 2 * L + u);
            b[2] = tw_get<PL, S_>(tw_s, tw_g, 3 * L + u);
            a[3] = make_float2(fmaf(-a[1].y, a[2].y, a[1].x * a[2].x), fmaf(a[1].x, a[2].y, a[1].y * a[2].x));
            b[3] = make_float2(fmaf(-b[1].y, b[2].y, b[1].x * b[2].x), fmaf(b[1].x, b[2].y, b[1].y * b[2].x));
        } else {
#pragma unroll
            for (int j = 1; j < 4; ++j) {
                a[j] = tw_get<PL, S_>(tw_s, tw_g, (j - 1) * L + u);
                b[j] = tw_get<PL, S_>(tw_s, tw_g, (j + 2) * L + u);
            }
        }
        apply_twiddles16(x, a, b);
    } else {
#pragma unroll
        for (int q = 1; q < R; ++q) {
            const float2 wq = tw_get<PL, S_>(tw_s, tw_g, (q - 1) * L + u);
            x[q] = cmul(x[q], wq.x, wq.y);
        }
    }
}
// ------------------------------------------------------------------ element traits per mode
template <int MODE>
struct Elem;
template <>
struct Elem<MODE_QUAD> {
    using V = float2;
    using S = float4;
    static constexpr int CH = 4;     // channels per element
    static constexpr int WAVE = 8;   // lanes per shared-memory wavefront for S
    static __device__ __forceinline__ Cx<V> unpack(const S &f) { return {make_float2(f.x, f.y), make_float2(f.z, f.w)}; }
    static __device__ __forceinline__ S pack(const Cx<V> &c) { return make_float4(c.re.x, c.re.y, c.im.x, c.im.y); }
};
template <>
struct Elem<MODE_PAIR> {
    using V = float;
    using S = float2;
    static constexpr int CH = 2;
    static constexpr int WAVE = 16;
    static __device__ __forceinline__ Cx<V> unpack(const S &f) { return {f.x, f.y}; }
    static __device__ __forceinline__ S pack(const Cx<V> &c) { return make_float2(c.re, c.im); }
};
template <>
struct Elem<MODE_REAL> {
    using V = float;
    using S = float2;
    static constexpr int CH = 1;
    static constexpr int WAVE = 16;
    static __device__ __forceinline__ Cx<V> unpack(const S &f) { return {f.x, f.y}; }
    static __device__ __forceinline__ S pack(const Cx<V> &c) { return make_float2(c.re, c.im); }
};

// global loads/stores of one element (CH channels of one row) --------------------------------
template <int MODE, class IO>
struct GIO;

template <>
struct GIO<MODE_QUAD, float> {
    static __device__ __forceinline__ Cx<float2> load(const float *p) {
        float4 f = __ldcs(reinterpret_cast<const float4 *>(p));
        return {make_float2(f.x, f.y), make_float2(f.z, f.w)};
    }
    static __device__ __forceinline__ void store(float *p, const Cx<float2> &c) {
        __stcs(reinterpret_cast<float4 *>(p), make_float4(c.re.x, c.re.y, c.im.x, c.im.y));
    }
};
template <>
struct GIO<MODE_QUAD, __nv_bfloat16> {
    static __device__ __forceinline__ Cx<float2> load(const __nv_bfloat16 *p) {
        uint2 r = __ldcs(reinterpret_cast<const uint2 *>(p));
        __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162 *>(&r.x), b = *reinterpret_cast<__nv_bfloat162 *>(&r.y);
        return {__bfloat1622float2(a), __bfloat1622float2(b)};
    }
    static __device__ __forceinline__ void store(__nv_bfloat16 *p, const Cx<float2> &c) {
        __nv_bfloat162 a = __float22bfloat162_rn(c.re), b = __float22bfloat162_rn(c.im);
        uint2 r;
        r.x = *reinterpret_cast<unsigned *>(&a);
        r.y = *reinterpret_cast<unsigned *>(&b);
        __stcs(reinterpret_cast<uint2 *>(p), r);
    }
};
template <>
struct GIO<MODE_PAIR, float> {
    static __device__ __forceinline__ Cx<float> load(const float *p) {
        float2 f = __ldcs(reinterpret_cast<const float2 *>(p));
        return {f.x, f.y};
    }
    static __device__ __forceinline__ void store(float *p, const Cx<float> &c) {
        __stcs(reinterpret_cast<float2 *>(p), make_float2(c.re, c.im));
    }
};
template <>
struct GIO<MODE_PAIR, __nv_bfloat16> {
    static __device__ __forceinline__ Cx<float> load(const __nv_bfloat16 *p) {
        unsigned r = __ldcs(reinterpret_cast<const unsigned *>(p));
        float2 f = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162 *>(&r));
        return {f.x, f.y};
    }
    static __device__ __forceinline__ void store(__nv_bfloat16 *p, const Cx<float> &c) {
        __nv_bfloat162 a = __float22bfloat162_rn(make_float2(c.re, c.im));
        __stcs(reinterpret_cast<unsigned *>(p), *reinterpret_cast<unsigned *>(&a));
    }
};
template <>
struct GIO<MODE_REAL, float> {
    static __device__ __forceinline__ Cx<float> load(const float *p) { return {__ldcs(p), 0.f}; }
    static __device__ __forceinline__ void store(float *p, const Cx<float> &c) { __stcs(p, c.re); }
};
template <>
struct GIO<MODE_REAL, __nv_bfloat16> {
    static __device__ __forceinline__ Cx<float> load(const __nv_bfloat16 *p) { return {__bfloat162float(*p), 0.f}; }
    static __device__ __forceinline__ void store(__nv_bfloat16 *p, const Cx<float> &c) { *p = __float2bfloat16_rn(c.re); }
};

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// ------------------------------------------------------------------ shared memory layout
template <class PL, int MODE, int NCOL>
struct Smem {
    using E = Elem<MODE>;
    static constexpr int SKEW = cmax(1, E::WAVE / NCOL);
    static constexpr int CS = PL::NPAD + SKEW;                 // column stride in elements
    static constexpr size_t data_bytes = sizeof(typename E::S) * (size_t)CS * NCOL;
    static constexpr size_t tw_bytes = ((sizeof(float2) * (size_t)TwPolicy<PL>::SMEM_N + 15) / 16) * 16;
    static constexpr size_t gate_bytes_one = ((sizeof(float2) * (size_t)PL::GPAD + 15) / 16) * 16;
    // TMA variant: 3 mbarriers + two output staging buffers of STG_MB stage-0 row blocks each
    __host__ __device__ static constexpr int cmin_(int a, int b) { return a < b ? a : b; }
    // stage-0 row blocks (L(0) rows each) per staging round: at least one TMA box (256 rows) when the tile has that many
    static constexpr int STG_MB = cmin_(PL::R(0), cmax(1, 256 / PL::L(0)) * (PL::N >= 4096 ? 2 : 1));
    static constexpr int STG_ROWS = STG_MB * PL::L(0);                       // rows per staging round
    static constexpr int OUT_BOX_ROWS = cmin_(STG_ROWS, 256);                // rows per TMA store
    __host__ __device__ static constexpr size_t stg_bytes(size_t row_bytes) { return (size_t)STG_ROWS * row_bytes; }
    static constexpr size_t bar_bytes = 256;
    __host__ __device__ static constexpr size_t base_bytes(int gate_tables) { return data_bytes + tw_bytes + gate_bytes_one * (size_t)gate_tables; }
    // TMEM variant: ring of kTmemSlots + kTmemStoreSlots TMA boxes (256 rows each) instead of the staging buffers
    static constexpr size_t bytes(int gate_tables, bool tma = false, size_t row_bytes = 0, bool tmem = false) {
        return ((base_bytes(gate_tables) + 127) / 128) * 128 + bar_bytes +
               (tmem ? (size_t)tmem_slots<PL>() * tmem_box_rows(NCOL) * row_bytes : (tma ? 2 * stg_bytes(row_bytes) : 0));
    }
};

// stage-s butterfly id -> frequency digits: k_low = sum_{i < NS-1} q_i P(i), given Q of the last stage
template <class PL>
__device__ __forceinline__ int klow_of(int Q) {
    // Q = ((q_0 R_1 + q_1) R_2 + q_2) ...  over stages 0 .. NS-2
    int k = 0;
#pragma unroll
    for (int s = PL::NS - 2; s >= 0; --s) {
        int q = Q % PL::R(s);
        Q /= PL::R(s);
        k += q * PL::P(s);
    }
    return k;
}

// ------------------------------------------------------------------ in-place inner passes (stages 1 .. NS-2)
// Butterfly bf = Q * L + u of a column works on elements Q*(R*L) + m*L + u, m < R, and leaves output
// digit q in the slot of input m = q, so no pass ever moves data between slots.
template <class PL, int MODE, int NCOL, int NT, int S_, bool ILV = false>
__device__ __forceinline__ void fwd_inner_pass(typename Elem<MODE>::S *buf, const float2 *tw, const float2 *twg, int tid) {
    using E = Elem<MODE>;
    using V = typename E::V;
    constexpr int R = PL::R(S_), L = PL::L(S_), NBF = PL::N / R, ITEMS = NCOL * NBF;
    constexpr int CS = Smem<PL, MODE, NCOL>::CS;
    // pad(e0 + m L) = pad(e0) + pad(m L) needs (e0 % 16) + (m L % 16) < 16: true for 16 | L, and for L | 16 because e0 % 16 = u < L
    static_assert(L % 16 == 0 || 16 % L == 0, "offset padding assumes 16 | L or L | 16");
    for (int w = tid; w < ITEMS; w += NT) {
        // ILV: the element columns are interleaved over adjacent lanes (both columns of a butterfly id share twiddles / gate)
        const int col = ILV ? w % NCOL : w / NBF, bf = ILV ? w / NCOL : w - col * NBF;
        const int Q = bf / L, u = bf - Q * L;
        const int e0 = Q * (R * L) + u;
        typename E::S *cb = buf + col * CS + e0 + (e0 >> 4);
        Cx<V> x[R];
#pragma unroll
        for (int m = 0; m < R; ++m) x[m] = E::unpack(cb[m * L + ((m * L) >> 4)]);
        Dft<R, V>::run(x);
        apply_twiddles<PL, S_, V>(x, tw, twg, u);
#pragma unroll
        for (int q = 0; q < R; ++q) cb[q * L + ((q * L) >> 4)] = E::pack(x[q]);
    }
}

// inverse of the above: conj-twiddle then inverse butterfly, done as forward arithmetic on (im, re)
template <class PL, int MODE, int NCOL, int NT, int S_, bool ILV = false>
__device__ __forceinline__ void inv_inner_pass(typename Elem<MODE>::S *buf, const float2 *tw, const float2 *twg, int tid) {
    using E = Elem<MODE>;
    using V = typename E::V;
    constexpr int R = PL::R(S_), L = PL::L(S_), NBF = PL::N / R, ITEMS = NCOL * NBF;
    constexpr int CS = Smem<PL, MODE, NCOL>::CS;
    for (int w = tid; w < ITEMS; w += NT) {
        // ILV: the element columns are interleaved over adjacent lanes (both columns of a butterfly id share twiddles / gate)
        const int col = ILV ? w % NCOL : w / NBF, bf = ILV ? w / NCOL : w - col * NBF;
        const int Q = bf / L, u = bf - Q * L;
        const int e0 = Q * (R * L) + u;
        typename E::S *cb = buf + col * CS + e0 + (e0 >> 4);
        Cx<V> x[R];
#pragma unroll
        for (int q = 0; q < R; ++q) x[q] = cswap(E::unpack(cb[q * L + ((q * L) >> 4)]));
        apply_twiddles<PL, S_, V>(x, tw, twg, u);
        Dft<R, V>::run(x);
#pragma unroll
        for (int m = 0; m < R; ++m) cb[m * L + ((m * L) >> 4)] = E::pack(cswap(x[m]));
    }
}

// spectral-memory add (spectre.py:548-549) on packed elements; sgn = +inv_n (bin k), -inv_n (mirror bin,
// conjugated) or 0 (imag of DC / Nyquist ignored, as irfft does)
__device__ __forceinline__ void mem_add(Cx<float2> &x, const float2 *mp, float inv_n, float sgn, int /*mode*/) {
    // channels c..c+3: z0 = ch0 + i ch2, z1 = ch1 + i ch3;  M_ch = a_ch + i b_ch
    const float4 m01 = __ldg(reinterpret_cast<const float4 *>(mp));
    const float4 m23 = __ldg(reinterpret_cast<const float4 *>(mp) + 1);
    x.re.x += m01.x * inv_n - m23.y * sgn;
    x.re.y += m01.z * inv_n - m23.w * sgn;
    x.im.x += m23.x * inv_n + m01.y * sgn;
    x.im.y += m23.z * inv_n + m01.w * sgn;
}
__device__ __forceinline__ void mem_add(Cx<float> &x, const float2 *mp, float inv_n, float sgn, int mode) {
    if (mode == MODE_PAIR) {  // z = ch0 + i ch1
        const float4 m01 = __ldg(reinterpret_cast<const float4 *>(mp));
        x.re += m01.x * inv_n - m01.w * sgn;
        x.im += m01.z * inv_n + m01.y * sgn;
    } else {                  // z = ch0
        const float2 m0 = __ldg(mp);
        x.re += m0.x * inv_n;
        x.im += m0.y * sgn;
    }
}

// ------------------------------------------------------------------ TMA / mbarrier primitives (sm_90+ PTX)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// box {channels, rows, 1} of the [B][rows][C] tensor -> dense [rows][channels] in shared memory
// SPX_L2HINT (experiment): L2 evict-first policy on the tile loads (bit 0) / the result stores (bit 1) -- every byte is touched once
#ifndef SPX_L2HINT
#define SPX_L2HINT 0
#endif
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void *tmap, int c, int row, int b, uint32_t bar) {
#if SPX_L2HINT & 1
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3, %4}], [%5], %6;"
        ::"r"(dst), "l"(tmap), "r"(c), "r"(row), "r"(b), "r"(bar), "l"(l2_evict_first_policy()) : "memory");
#else
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(dst), "l"(tmap), "r"(c), "r"(row), "r"(b), "r"(bar) : "memory");
#endif
}
__device__ __forceinline__ void tma_prefetch_3d(const void *tmap, int c, int row, int b) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(tmap), "r"(c), "r"(row), "r"(b)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// dense [rows][channels] in shared memory -> box {channels, rows, 1} of the [B][rows][C] output (clipped at the edges)
__device__ __forceinline__ void tma_store_3d(const void *tmap, uint32_t src, int c, int row, int b) {
#if SPX_L2HINT & 2
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group.L2::cache_hint [%0, {%1, %2, %3}], [%4], %5;" ::"l"(tmap), "r"(c),
                 "r"(row), "r"(b), "r"(src), "l"(l2_evict_first_policy())
                 : "memory");
#else
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(tmap), "r"(c),
                 "r"(row), "r"(b), "r"(src)
                 : "memory");
#endif
}
// rank-4 forms for the DIT2 variant: tensor [B][n][parity][C] (row = 2 n + parity), box {channels, 2, rows, 1}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const void *tmap, int c, int par, int row, int b, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(dst), "l"(tmap), "r"(c), "r"(par), "r"(row), "r"(b), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_4d(const void *tmap, uint32_t src, int c, int par, int row, int b) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3, %4}], [%5];" ::"l"(tmap), "r"(c),
                 "r"(par), "r"(row), "r"(b), "r"(src)
                 : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void tma_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(PENDING) : "memory");
}
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------ tensor memory (TMEM) as an I/O staging store
// 512 columns x 128 lanes x 32 bit per SM.  A warp reaches the 32 lanes of quadrant (warp id % 4); thread i of the warp is
// lane base + i; address = lane << 16 | column.  Measured here: 128 KB written in 127 ns, read in 174 ns per SM
// (tools/microbench/tmem_probe.cu) -- 3-4x shared memory -- so parking a tile costs next to nothing.
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t a, float &r0, float &r1, float &r2, float &r3) {
    uint32_t u0, u1, u2, u3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(u0), "=r"(u1), "=r"(u2), "=r"(u3) : "r"(a) : "memory");
    r0 = __uint_as_float(u0); r1 = __uint_as_float(u1); r2 = __uint_as_float(u2); r3 = __uint_as_float(u3);
}
__device__ __forceinline__ void tmem_st4(uint32_t a, float r0, float r1, float r2, float r3) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(__float_as_uint(r0)),
                 "r"(__float_as_uint(r1)), "r"(__float_as_uint(r2)), "r"(__float_as_uint(r3)) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- warp-local 16 x 16 transposes THROUGH tensor memory (no shared-memory traffic; tools/microbench/tmem_xchg.cu)
// A .32x32b store puts register column c of lane s at (lane s, column c); a .16x256b load at lane half hh hands thread t the
// columns 8 i + 2 (t % 4) + {0, 1} of lanes t / 4 + 8 v + 16 hh.  One store + two loads therefore move the top two lane bits of
// the source into the register index and two register-index bits into the low two lane bits, shifting the other lane bits up by
// two.  Two such steps turn the stage-1 layout (lane = 2 u + col, register q) into the middle-pass layout (lane = 16 col + q,
// register u); the twin with the shapes swapped is the exact inverse.  Each step runs as four 16-column quarters, so a warp needs
// 16 columns of tensor memory.  LDTM / STTM run on their own pipe: measured 1400 cycles per 128 KB exchange and SM, and almost
// fully hidden behind shared-memory traffic, against 2050 cycles of LSU time for the shared-memory exchange.
__device__ __forceinline__ void tmem_st_32x16(uint32_t a, const float (&r)[16]) {
#define SPX_U(i) "r"(__float_as_uint(r[i]))
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(a),
                 SPX_U(0), SPX_U(1), SPX_U(2), SPX_U(3), SPX_U(4), SPX_U(5), SPX_U(6), SPX_U(7), SPX_U(8), SPX_U(9), SPX_U(10), SPX_U(11),
                 SPX_U(12), SPX_U(13), SPX_U(14), SPX_U(15) : "memory");
}
__device__ __forceinline__ void tmem_st_16x256x2(uint32_t a, const float (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(a), SPX_U(0), SPX_U(1), SPX_U(2),
                 SPX_U(3), SPX_U(4), SPX_U(5), SPX_U(6), SPX_U(7) : "memory");
#undef SPX_U
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t a, float (&r)[16]) {
    uint32_t u[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
                   "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]) : "r"(a) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = __uint_as_float(u[i]);
}
__device__ __forceinline__ void tmem_ld_16x256x2(uint32_t a, float (&r)[8]) {
    uint32_t u[8];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]) : "r"(a) : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = __uint_as_float(u[i]);
}
// word f of a packed element: (re.x, re.y, im.x, im.y) -- the two halves of a packed-fp32x2 register pair stay adjacent columns
__device__ __forceinline__ float &cx_word(Cx<float2> &c, int f) { return f == 0 ? c.re.x : (f == 1 ? c.re.y : (f == 2 ? c.im.x : c.im.y)); }
// HI: exchange element-index bits 3:2 (quarters = bits 1:0); else bits 1:0 (quarters = bits 3:2)
template <bool HI>
__device__ __forceinline__ void tmem_xchg_step_fwd(Cx<float2> (&x)[16], uint32_t tx) {
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        float s[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) s[j] = cx_word(x[HI ? 4 * ((j >> 1) & 3) + h : 4 * h + ((j >> 1) & 3)], (j & 1) + 2 * (j >> 3));
        tmem_st_32x16(tx, s);
        tmem_wait_st();
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            float r[8];   // r[4 i + 2 v + f0]: repetition i = word bit 1, v = source lane + 8, f0 = word bit 0
            tmem_ld_16x256x2(tx + ((uint32_t)(16 * hh) << 16), r);
#pragma unroll
            for (int j = 0; j < 8; ++j) cx_word(x[HI ? 4 * (2 * hh + ((j >> 1) & 1)) + h : 4 * h + 2 * hh + ((j >> 1) & 1)], (j & 1) + 2 * (j >> 2)) = r[j];
        }
        tmem_wait_ld();
    }
}
template <bool HI>
__device__ __forceinline__ void tmem_xchg_step_inv(Cx<float2> (&x)[16], uint32_t tx) {
#pragma unroll
    for (int h = 0; h < 4; ++h) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            float r[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = cx_word(x[HI ? 4 * (2 * hh + ((j >> 1) & 1)) + h : 4 * h + 2 * hh + ((j >> 1) & 1)], (j & 1) + 2 * (j >> 2));
            tmem_st_16x256x2(tx + ((uint32_t)(16 * hh) << 16), r);
        }
        tmem_wait_st();
        float s[16];
        tmem_ld_32x16(tx, s);
        tmem_wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) cx_word(x[HI ? 4 * ((j >> 1) & 3) + h : 4 * h + ((j >> 1) & 3)], (j & 1) + 2 * (j >> 3)) = s[j];
    }
}
// stage-1 layout -> middle-pass layout and back (see above)
__device__ __forceinline__ void tmem_xchg_fwd(Cx<float2> (&x)[16], uint32_t tx) { tmem_xchg_step_fwd<true>(x, tx); tmem_xchg_step_fwd<false>(x, tx); }
__device__ __forceinline__ void tmem_xchg_inv(Cx<float2> (&x)[16], uint32_t tx) { tmem_xchg_step_inv<false>(x, tx); tmem_xchg_step_inv<true>(x, tx); }
template <class T> __device__ __forceinline__ void tmem_xchg_fwd(T &, uint32_t) {}   // other element types: never instantiated for use
template <class T> __device__ __forceinline__ void tmem_xchg_inv(T &, uint32_t) {}
template <bool FIRST, class A, class B>
__device__ __forceinline__ auto &pick_ref(A &a, B &b) {
    if constexpr (FIRST) return a; else return b;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void helper_bar() { asm volatile("bar.sync 2, 128;" ::: "memory"); }
// Programmatic dependent launch (the host side sets the launch attribute unless sched bit 7 is set): everything above the wait -- barrier
// set-up, tensor-memory allocation, the twiddle table (a constant built once per device) -- may run while the previous kernel of
// the stream is still draining; nothing produced by that kernel is read, and nothing it may still read is written, before the
// wait.  The trigger right behind it lets the NEXT launch's CTAs take an SM as soon as one of ours exits.  Both are no-ops
// for a launch without the attribute.
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// Warp stagger: the four warps that share a scheduler (warp id / 4 = "slot" 0..3) leave a CTA barrier in lock step and would
// all load, then all compute, then all store.  Holding slot s back by s * clk cycles lets the shared-memory phase of one warp
// run under the butterflies of another.  code > 0: legacy nanosleep of slots 1 and 3; code < 0: clock spin, |code| % 100000
// cycles per step, |code| / 100000 selects the grouping (0: four steps, 1: slots {0,1} vs {2,3}, 2: even vs odd slots,
// 3: four steps plus a quarter step per scheduler, 4: four steps by __nanosleep instead of a clock spin).
__device__ __forceinline__ void stagger(int code, int tid) {
    if (code == 0) return;
    if (code > 0) {
        if ((tid >> 7) & 1) __nanosleep(code);
        return;
    }
    const int a = -code, grp = a / 100000, clk = a % 100000;
    int slot = (tid >> 7) & 3;
    if (grp == 1) slot >>= 1;
    else if (grp == 2) slot &= 1;
    int target = slot * clk;
    if (grp == 3) target += ((tid >> 5) & 3) * (clk >> 2);   // 16 distinct offsets: the schedulers' warps a quarter step apart
    if (grp == 4) {                                            // sleep instead of spinning (no issue slots while waiting)
        if (target > 0) __nanosleep((unsigned)(target / 2));   // ~2 clocks per ns
        return;
    }
    const int t0 = (int)clock();
    while ((int)clock() - t0 < target) {}
}

// landing-buffer element (one tile element as TMA delivers it: CH consecutive channels of one row)
template <int MODE, class IO>
struct Lin;
template <>
struct Lin<MODE_QUAD, float> {
    using T = float4;
    static __device__ __forceinline__ Cx<float2> get(const T &f) { return {make_float2(f.x, f.y), make_float2(f.z, f.w)}; }
    static __device__ __forceinline__ T put(const Cx<float2> &c) { return make_float4(c.re.x, c.re.y, c.im.x, c.im.y); }
};
template <>
struct Lin<MODE_QUAD, __nv_bfloat16> {
    using T = uint2;
    static __device__ __forceinline__ Cx<float2> get(const T &r) {
        __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162 *>(&r.x), b = *reinterpret_cast<const __nv_bfloat162 *>(&r.y);
        return {__bfloat1622float2(a), __bfloat1622float2(b)};
    }
    static __device__ __forceinline__ T put(const Cx<float2> &c) {
        __nv_bfloat162 a = __float22bfloat162_rn(c.re), b = __float22bfloat162_rn(c.im);
        uint2 r;
        r.x = *reinterpret_cast<unsigned *>(&a);
        r.y = *reinterpret_cast<unsigned *>(&b);
        return r;
    }
};
template <int MODE, class IO>
struct Lin {  // PAIR / REAL never use the TMA path
    using T = float2;
    static __device__ __forceinline__ Cx<float> get(const T &f) { return {f.x, f.y}; }
    static __device__ __forceinline__ T put(const Cx<float> &c) { return make_float2(c.re, c.im); }
};


// The TMA variant adds one producer warpgroup (one elected lane works).  A whole warpgroup, because setmaxnreg moves
// registers between warpgroups: the producer shrinks to 24 registers and the compute warps take what it frees.
constexpr int kProducerThreads = 128;
// CTAs with fewer compute threads keep the producer duties on compute thread 0 (an extra warpgroup would cost them
// occupancy and registers)
constexpr int kSepProducerMinThreads = 256;
__host__ __device__ constexpr int helper_regs(int nt, int minb);
__host__ __device__ constexpr int reg_floor8(int r) { return r / 8 * 8; }
// register budget of the compute warps after the hand-over (launch allocation = nominal * all threads)
// what is left per helper thread once the compute warps took their share
__host__ __device__ constexpr int helper_regs_raw(int nt, int minb, int creg) {
    return reg_floor8((reg_floor8(65536 / (minb * (nt + kProducerThreads))) * (nt + kProducerThreads) - creg * nt) / kProducerThreads);
}
__host__ __device__ constexpr int compute_regs(int nt, int minb) {
    return reg_floor8((reg_floor8(65536 / (minb * (nt + kProducerThreads))) * (nt + kProducerThreads) - 24 * kProducerThreads) / nt) > 128
               ? 128
               : reg_floor8((reg_floor8(65536 / (minb * (nt + kProducerThreads))) * (nt + kProducerThreads) - 24 * kProducerThreads) / nt);
}
__host__ __device__ constexpr int helper_regs(int nt, int minb) {
    return helper_regs_raw(nt, minb, compute_regs(nt, minb)) < 24 ? 24 : helper_regs_raw(nt, minb, compute_regs(nt, minb));
}

// One gate row (batch row b, gate row g): a pointer into the materialised gate tensor, or (ANCH) the anchors of the row's
// head from which spx::gate_from_anchors evaluates a bin on demand (spectre.py:526-536 fused into the gate staging).
template <bool ANCH>
struct GateRow {
    const float2 *row;     // !ANCH: gate + (b NG + g) F_half
    const float2 *a;       // ANCH: anchors of the head, [G][Bk]
    const float *bias;     // ANCH: modReLU bias of this gate row, [F_half]
    const float2 *pos;     // ANCH: positional phase row or nullptr
    const float4 *icoef;   // ANCH: interpolation table (or nullptr)
    const ushort4 *itap;
    float eps;
    int Bk, G, j, F_half;
    __device__ __forceinline__ float2 at(int k) const {
        if constexpr (ANCH) return gate_from_anchors_fast(icoef, itap, Bk, G, a, j, k, F_half, __ldg(bias + k), eps, pos);
        else return __ldg(row + k);
    }
};
template <bool ANCH>
__device__ __forceinline__ GateRow<ANCH> gate_row(const MixParams &p, int b, int g, int F_half) {
    GateRow<ANCH> r;
    if constexpr (ANCH) {
        const GateSrc &s = p.gsrc;
        const int head = g / s.G;
        r.row = nullptr;
        r.a = s.anchors + ((size_t)b * p.NG + (size_t)head * s.G) * s.Bk;
        r.bias = s.bias + (size_t)g * F_half;
        r.pos = s.pos ? s.pos + (size_t)b * s.pos_stride_b : nullptr;
        r.eps = __ldg(s.eps + g);
        r.icoef = s.icoef;
        r.itap = s.itap;
        r.Bk = s.Bk;
        r.G = s.G;
        r.j = g - head * s.G;
        r.F_half = F_half;
    } else {
        r.row = p.gate + ((long long)b * p.NG + g) * F_half;
        r.a = nullptr; r.bias = nullptr; r.pos = nullptr; r.icoef = nullptr; r.itap = nullptr; r.eps = 0.f; r.Bk = 0; r.G = 1; r.j = 0; r.F_half = F_half;
    }
    return r;
}

// gate row -> registers (all loads in flight), registers -> padded shared table; J0 = first of the thread's entries (k = tid + j NT)
template <int N, int NT, int GK, bool ANCH, int J0 = 0>
__device__ __forceinline__ void gate_fetch(float2 (&gv)[GK], const GateRow<ANCH> &gr, int tid) {
#pragma unroll
    for (int j = 0; j < GK; ++j) {
        const int k = tid + (j + J0) * NT;
        gv[j] = (k <= N / 2) ? gr.at(k) : make_float2(0.f, 0.f);
    }
}
template <int N, int NT, int GK, int J0 = 0>
__device__ __forceinline__ void gate_put(float2 *gs, const float2 (&gv)[GK], int tid, float inv_n) {
#pragma unroll
    for (int j = 0; j < GK; ++j) {
        const int k = tid + (j + J0) * NT;
        if (k <= N / 2) {
            const float im = (k == 0 || k == N / 2) ? 0.f : gv[j].y * inv_n;
            gs[k + (k >> 4)] = make_float2(gv[j].x * inv_n, im);
        }
    }
}

// gate row -> padded shared table by asynchronous 8-byte copies (LDGSTS): no registers are held while the row is in flight.
// Once its copies have landed each thread rescales its own entries in place (1/n_fft, imag(DC) = imag(Nyquist) = 0).
template <int N, int NT, int GK>
__device__ __forceinline__ void gate_copy_async(float2 *gs, const float2 *gp, int tid) {
#pragma unroll
    for (int j = 0; j < GK; ++j) {
        const int k = tid + j * NT;
        if (k <= N / 2)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(gs + k + (k >> 4))), "l"(gp + k) : "memory");
    }
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
template <int N, int NT, int GK>
__device__ __forceinline__ void gate_scale_own(float2 *gs, int tid, float inv_n) {
#pragma unroll
    for (int j = 0; j < GK; ++j) {
        const int k = tid + j * NT;
        if (k <= N / 2) {
            const float2 g = gs[k + (k >> 4)];
            gs[k + (k >> 4)] = make_float2(g.x * inv_n, (k == 0 || k == N / 2) ? 0.f : g.y * inv_n);
        }
    }
}

// the same two steps for the helper warpgroup (128 threads at 32 registers: rolled loops, nothing hoisted)
template <int N>
__device__ __forceinline__ void gate_copy_async_rolled(float2 *gs, const float2 *gp, int t) {
#pragma unroll 1
    for (int k = t; k <= N / 2; k += 128)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(gs + k + (k >> 4))), "l"(gp + k) : "memory");
}
template <int N>
__device__ __forceinline__ void gate_scale_own_rolled(float2 *gs, int t, float inv_n) {
#pragma unroll 1
    for (int k = t; k <= N / 2; k += 128) {
        const float2 g = gs[k + (k >> 4)];
        gs[k + (k >> 4)] = make_float2(g.x * inv_n, (k == 0 || k == N / 2) ? 0.f : g.y * inv_n);
    }
}

// barrier over the NT compute threads only (the TMA producer warp of the TMA variant never joins it)
template <int NT, bool NAMED>
__device__ __forceinline__ void cta_sync() {
    if constexpr (NAMED) asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
    else __syncthreads();
}

// sub-transform q of a length n_total = R * N transform: bin k' of the sub-spectrum is bin k = q + R k' of the long one;
// table[k'] = Gfull[k] / n_total with Gfull the Hermitian extension (imag of DC / Nyquist dropped)
template <int N, int NT, int GKS, bool ANCH>
__device__ __forceinline__ void gate_fetch_sub(float2 (&gv)[GKS], const GateRow<ANCH> &gr, int tid, int q, int R) {
    const int nt = N * R;
#pragma unroll
    for (int j = 0; j < GKS; ++j) {
        const int k = q + R * (tid + j * NT);
        gv[j] = gr.at(k <= nt / 2 ? k : nt - k);
    }
}
template <int N, int NT, int GKS>
__device__ __forceinline__ void gate_put_sub(float2 *gs, const float2 (&gv)[GKS], int tid, int q, int R, float inv_n) {
    const int nt = N * R;
#pragma unroll
    for (int j = 0; j < GKS; ++j) {
        const int kp = tid + j * NT;
        const int k = q + R * kp;
        float im = (k <= nt / 2) ? gv[j].y : -gv[j].y;
        if (k == 0 || k == nt / 2) im = 0.f;
        gs[kp + (kp >> 4)] = make_float2(gv[j].x * inv_n, im * inv_n);
    }
}

// ------------------------------------------------------------------ the kernel
// One persistent CTA per resident slot; each loop iteration transforms one tile = all n_fft rows of NCOL
// elements (CH channels each) of one batch row.  RFFT_ONLY: stop after the forward half and write the half
// spectrum.  TMA_IN: the tile is brought into shared memory by TMA (cp.async.bulk.tensor) and the CTA's next
// tile is prefetched into L2 by the TMA unit; otherwise stage 0 loads straight from global into registers.
template <class PL, int MODE, int NCOL, int NT, int MINB, class TIN, class TOUT, bool HAS_MEM, bool RFFT_ONLY = false,
          bool TMA_IN = false, bool TMEM_IO = false, bool ANCH = false, bool DGATE = false>
__global__ void __launch_bounds__(NT + ((TMA_IN && NT >= kSepProducerMinThreads) ? kProducerThreads : 0), MINB)
    spectre_mix_kernel(const MixParams p, const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_out) {
    using E = Elem<MODE>;
    using V = typename E::V;
    using S = typename E::S;
    using SM = Smem<PL, MODE, NCOL>;
    constexpr int N = PL::N, NS = PL::NS, CH = E::CH, CS = SM::CS;
    constexpr int RL = PL::R(NS - 1);       // radix of the last stage (fused middle pass)
    constexpr int PLAST = PL::P(NS - 1);    // = N / RL
    constexpr int R0 = PL::R(0), L0 = PL::L(0);
    constexpr int ITEMS0 = NCOL * L0;                       // stage-0 butterflies per tile
    constexpr int ITERS0 = (ITEMS0 + NT - 1) / NT;          // per thread
    static_assert(!TMA_IN || (MODE == MODE_QUAD && !RFFT_ONLY), "TMA path is built for the packed mix kernel");
    static_assert(!ANCH || !RFFT_ONLY, "in-kernel gate generation: mix kernels");
    // DGATE (backward of the mix w.r.t. the gate): every tile is visited twice -- first with V's rows (forward
    // transform, packed spectrum stashed in this thread's own tensor-memory lane: the TMEM-OUT half is free, nothing is
    // drained), then with dY's rows (tensor map 2), whose spectrum is multiplied by the conjugate of the stash, reduced over the
    // channels of the gate group and over the mirrored bin, and added to dgate[b, g, :].  No inverse passes, no gate, no output tile.
    static_assert(!DGATE || (TMEM_IO && !RFFT_ONLY && !ANCH && !HAS_MEM && !PL::kSub && MODE == MODE_QUAD && PL::NS == 3),
                  "dgate kernel: TMEM-staged packed variant of a three-stage plan");
    // stage NS-2 and the middle pass both have radix 16 and >= 32 butterflies per column, items map to threads
    // identically in both (w = tid + k NT): their exchange stays inside a warp
    constexpr bool kWarpLocal = (NS >= 3) && (PL::R(NS - 1) == 16) && (PL::R(NS - 2) == 16) && (NT % 32 == 0) &&
                                ((NCOL * (N / 16)) % NT == 0 || NT % (NCOL * (N / 16)) == 0) && ((N / 16) % 32 == 0);

    // Narrow last radix (wide-row TMEM variants: 1024 = 16 x 16 x 4 with 32-channel tiles, 2048 = 16 x 16 x 8 with 16-channel
    // tiles): stage 1 has L = RL, so the RL threads (same leading digits, u = 0 .. RL-1; adjacent lanes) of a stage-1 butterfly
    // group produce exactly the inputs of 16 middle-pass items; giving each of them 16 / RL of those items keeps the exchange
    // inside the warp here too (see the middle pass).
    constexpr bool kWarpLocalN = TMEM_IO && (NS == 3) && (PL::R(1) == 16) && (RL < 16) && (16 % RL == 0) && (NCOL * (N / 16) == NT);
    // inner passes and the middle pass map (column, butterfly) -> thread with the two element columns on adjacent lanes: gate and
    // twiddle reads of a lane pair coincide (one wavefront instead of two); the exchanges stay inside a warp
    constexpr bool kIlv = (SPX_ILV != 0) && TMEM_IO && kWarpLocal && NS == 3 && NCOL == 2 && MODE == MODE_QUAD;
    // stage 1, the middle pass and inverse stage 1 in registers, linked by two warp-local transposes through tensor memory
    // (tmem_xchg_fwd / _inv): one thread = one stage-1 butterfly = one middle-pass item.  The 16 columns a warp needs are cells of
    // TMEM-IN that belong to the LAST kTmemXBoxes input boxes: free once every warp has pulled its tile (the barrier after stage
    // 0), refilled by the helper only after every warp has reported its second exchange done (barrier +120).
    constexpr bool kTmemX = ((SPX_TMEMX & 1) != 0) && kIlv && !DGATE && !PL::kDit && !PL::kSub && !RFFT_ONLY && (NCOL * (N / 16) == NT) && (NT == 512);
    constexpr int kTmemXBoxes = 4;
    // the way back (middle pass -> inverse stage 1) through tensor memory as well.  Off by default: it keeps the 16 elements live
    // across the next tile's gate prefetch, and ptxas then spills the freshly loaded gate row (a wait for DRAM in every tile);
    // with the row sent by LDGSTS instead (SPX_GATE_ASYNC) the tile start pays.  Measured 10.0 / 11.1 us per tile against 9.2.
    constexpr bool kTmemXI = kTmemX && ((SPX_TMEMX & 2) != 0);
    // The next tile's gate row goes into the table by asynchronous 8-byte copies instead of through registers.  With kTmemX the
    // registers that would park the row across inverse stage 1 do not exist (the transform's 16 elements stay live from stage 1
    // to the last inverse pass): ptxas spilled the freshly loaded row to local memory, i.e. waited for DRAM right there.
    constexpr bool kGateAsync = (SPX_GATE_ASYNC == 1 || SPX_GATE_ASYNC == 3 || (SPX_GATE_ASYNC == 2 && kTmemXI)) && !ANCH && !PL::kSub &&
                                !PL::kDit && !DGATE && !RFFT_ONLY && (SPX_GATE_ASYNC != 3 || (TMEM_IO && NCOL * CH < 32));
    // mode 3: the table keeps the RAW row (no rescaling pass); the middle pass applies 1/n and the imag(DC) = imag(Nyquist) = 0
    // rule where it reads an entry, as the DIT2 kernel does
    constexpr bool kGateRaw = kGateAsync && (SPX_GATE_ASYNC == 3);
    // Experiment (off): the helper warpgroup stages the gate rows (LDGSTS into the table at the start of the phase that follows the
    // tile's parking, its own entries rescaled two steps later, published on barrier +128).  Without any gate staging the compute
    // warps run 9.0 instead of 9.2 us per tile (profiles/r02x_ab_nogate.txt), but the helper's loop is nearly as long as the
    // compute warps' and every wait added to it stalls the load stream: 10.4 - 10.8 us per tile with the row staged there.
    // DIT2 (n_fft = 8192): the gate row is 33 KB, nine entries per thread.  Parked in registers across inverse stage 1 it gets
    // spilled by ptxas right behind its loads (a wait for DRAM in every tile); so the row goes into the table by asynchronous
    // 8-byte copies after the barrier that ends inverse stage 1, raw, and the middle pass applies 1/n and the imag(DC) =
    // imag(Nyquist) = 0 rule when it reads an entry (four scalar products per bin pair).
    constexpr bool kDitAsync = (SPX_DIT_ASYNC != 0) && PL::kDit && !ANCH && !RFFT_ONLY && !DGATE;
    constexpr bool kHelperGate = (SPX_HELPER_GATE != 0) && TMEM_IO && !ANCH && !PL::kSub && !PL::kDit && !DGATE && !RFFT_ONLY && !kGateAsync;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    S *buf = reinterpret_cast<S *>(smem_raw);
    float2 *tw = reinterpret_cast<float2 *>(smem_raw + SM::data_bytes);
    float2 *gate_s = reinterpret_cast<float2 *>(smem_raw + SM::data_bytes + SM::tw_bytes);
    constexpr int GS = (int)(SM::gate_bytes_one / sizeof(float2));
    // mbarriers and the output staging buffers live after the gate tables
    //   +0 landed (TMA tx)   +8 buffer free (NW warps)   +16/+24 staging full (NW warps)   +32/+40 staging free (producer)
    const size_t bar_off = ((SM::base_bytes(p.gate_tables) + 127) / 128) * 128;
    const uint32_t bar = smem_u32(smem_raw + bar_off);
    const uint32_t bar_buf_free = bar + 8, bar_stg_full = bar + 16, bar_stg_free = bar + 32;
    unsigned char *stg = smem_raw + bar_off + SM::bar_bytes;
    constexpr int NW = NT / 32;
    constexpr bool SEP = TMA_IN && (NT >= kSepProducerMinThreads);   // separate producer warpgroup
    [[maybe_unused]] constexpr int kTmemSlots = tmem_slots<PL>();        // ring slots of the TMEM variant
    // TMEM layout of a parked tile: the tile as a flat array of elements e = row * NCOL + col (16 bytes each), element e in
    // lane e % 128, columns 4 * (e / 128) .. +3.  Helper threads then touch consecutive 16-byte slots of the ring (no bank
    // conflicts) and a compute thread finds all 16 inputs of its stage-0 butterfly in its own lane.
    // TMEM_IO: the helper warpgroup parks the NEXT tile in tensor memory while this one is transformed and drains the
    // PREVIOUS tile's results from tensor memory, so loads, stores and the FFT passes of three tiles overlap although
    // shared memory holds only one.  Needs one stage-0 butterfly per thread and row blocks that are multiples of 128.
    static_assert(!TMEM_IO || (SEP && sizeof(TIN) == sizeof(TOUT) && NT == NCOL * PL::L(0) && (PL::L(0) * NCOL) % 128 == 0 &&
                               128 % NCOL == 0 && MINB == 1 && (PL::N * NCOL * 4) % 128 == 0 && PL::N * NCOL * 4 / 128 * 2 <= 512),
                  "TMEM staging: unsupported shape");
    // tensor-memory columns between the rows u + L(0) m and u + L(0) (m + 1) of a stage-0 butterfly: the flat element index
    // advances by L(0) * NCOL, a multiple of 128, i.e. by L(0) * NCOL / 128 column groups of 4
    constexpr int MSTRIDE = 4 * (PL::L(0) * NCOL / 128);
    constexpr int CPR = NCOL * 4;                        // TMEM columns per tile row (fp32 channels)
    constexpr int TCOLS = PL::N / 128 * CPR;             // TMEM columns of one parked tile (256 at 4096 x 8 ch)
    const int tid = threadIdx.x;
    // twiddle tables -> shared memory, once per (persistent) CTA: one bulk-async copy (TMA, UBLKCP) in the TMA variants,
    // plain loads otherwise; the first barrier of the tile loop publishes the table
    constexpr uint32_t TW_BYTES = (uint32_t)(sizeof(float2) * TwPolicy<PL>::SMEM_N);
    constexpr bool TW_BULK = TMA_IN && (TW_BYTES % 16 == 0) && TW_BYTES > 0;
    if constexpr (!TW_BULK) {
        for (int i = tid; i < TwPolicy<PL>::SMEM_N; i += NT) tw[i] = p.tw[PL::TWOFF(TwPolicy<PL>::FROM) + i];
    }

    constexpr int GK = (N / 2 + 1 + NT - 1) / NT;   // gate entries per thread
    constexpr bool SUB = PL::kSub;
    constexpr bool DIT = PL::kDit;                  // columns = even / odd rows of a 2 N-point transform (see Plan)
    constexpr int NTOT = DIT ? 2 * N : N;           // transform length the gate / memory belong to
    constexpr int GKD = (NTOT / 2 + 1 + NT - 1) / NT;   // DIT: gate entries per thread (half spectrum of the 2 N-point transform)
    constexpr int GKD1 = DIT ? 5 : 1;               // ... of which this many are parked in registers across the inner inverse pass,
    constexpr int GKD2 = DIT ? GKD - GKD1 : 1;      // the rest is fetched and published right after it (register pressure)
    static_assert(!DIT || (TMEM_IO && NCOL == 2 && MODE == MODE_QUAD && !RFFT_ONLY && !DGATE && PL::NS == 3 && PL::R(2) == 16),
                  "DIT2 variant: TMEM-staged packed kernel with two element columns");
    constexpr int GKS = SUB ? N / NT : 1;           // sub-transform: full-length gate table, entries per thread
    static_assert(!SUB || (N % NT == 0 && MODE == MODE_QUAD && !RFFT_ONLY), "sub-transform variant: packed mix kernel only");
    // tables fetched for the NEXT tile while this one finishes (parked in registers across the inner inverse pass): one, or
    // two on the wide tiles whose 32 channels span two 16-channel gate groups
    constexpr int kEarlyGT = (!SUB && NCOL * CH >= 32) ? 2 : 1;
    const bool gate_early = SUB || DIT || (p.gate_tables <= kEarlyGT);
    const bool hgate = kHelperGate && p.gate_tables == 1;   // the helper warpgroup stages the gate rows (see kHelperGate)
    const TIN *vbase = reinterpret_cast<const TIN *>(p.v);
    TOUT *obase = reinterpret_cast<TOUT *>(p.out);
    const int CE = p.C / CH;  // elements per row

    using LT = typename Lin<MODE, TIN>::T;
    constexpr int BOXR = N < kTmaBoxRows ? N : kTmaBoxRows;
    constexpr uint32_t ROWB = (uint32_t)(sizeof(LT) * NCOL);        // landed bytes per row
    auto issue_tile_load = [&](int t) {                              // one thread
        const int tb = t / p.tiles_per_row;
        const int tc = (t - tb * p.tiles_per_row) * NCOL * CH;
        mbar_expect_tx(bar, ROWB * N);
#pragma unroll 1
        for (int r0 = 0; r0 < N; r0 += BOXR) tma_load_3d(smem_u32(smem_raw) + r0 * ROWB, &tmap, tc, r0, tb, bar);
    };
    auto prefetch_tile = [&](int t) {                                // one thread; rows past n_in need no traffic
        const int tb = t / p.tiles_per_row;
        const int tc = (t - tb * p.tiles_per_row) * NCOL * CH;
#pragma unroll 1
        for (int r0 = 0; r0 < p.n_in; r0 += BOXR) tma_prefetch_3d(&tmap, tc, r0, tb);
    };
    using OTs = typename Lin<MODE, TOUT>::T;
    constexpr uint32_t OROWB = (uint32_t)(sizeof(OTs) * NCOL);
    constexpr uint32_t STGB = (uint32_t)SM::stg_bytes(OROWB);
    constexpr int NR = PL::R(0) / SM::STG_MB;                          // output staging rounds per tile
    // producer duties (one thread): next-tile load once the buffer is free; TMA stores of a filled staging buffer
    uint32_t rnd_p = 0;
    auto producer_next_load = [&](int tile, int it) {
        const int nt = tile + gridDim.x;
        mbar_wait(bar_buf_free, it & 1);       // every compute warp has pulled `tile` out of the buffer
        if (nt < p.num_tiles) issue_tile_load(nt);
    };
    auto producer_store_round = [&](int tile, int j) {
        const int tb = tile / p.tiles_per_row;
        const int tc = (tile - tb * p.tiles_per_row) * NCOL * CH;
        const uint32_t kbuf = rnd_p & 1;
        mbar_wait(bar_stg_full + 8 * kbuf, (rnd_p >> 1) & 1);
        const uint32_t sb = smem_u32(stg + kbuf * STGB);
#pragma unroll 1
        for (int r0 = 0; r0 < SM::STG_ROWS; r0 += SM::OUT_BOX_ROWS)
            tma_store_3d(&tmap_out, sb + r0 * OROWB, tc, j * SM::STG_ROWS + r0, tb);
        tma_commit();
        tma_wait_read<1>();                    // every store but the one just issued has left shared memory
        if (rnd_p >= 1) mbar_arrive(bar_stg_free + 8 * (kbuf ^ 1));
        ++rnd_p;
    };
    if constexpr (TW_BULK) {
        const uint32_t bar_tw = bar + 104;
        if (tid == 0) {
            mbar_init(bar_tw, 1);
            mbar_expect_tx(bar_tw, TW_BYTES);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(tw)),
                         "l"(p.tw + PL::TWOFF(TwPolicy<PL>::FROM)), "r"(TW_BYTES), "r"(bar_tw)
                         : "memory");
            mbar_wait(bar_tw, 0);   // the set-up barrier below (all threads) orders this before any use
        }
    }
    // ---------------------------------------------------------------- TMEM-staged variant: set-up + helper warpgroup
    // barriers:  +0 tile parked in TMEM-IN (helper)   +8 TMEM-IN consumed (NW warps)   +16 results parked in TMEM-OUT (NW warps)
    //            +24 TMEM-OUT drained (helper)   +32.. ring slot landed (TMA tx, kTmemSlots of them, < +96)   +96 TMEM base address
    //            +104 twiddle table landed   +112 last inverse pass has read the buffer (NW warps, split barrier)
    //            +120 both tensor-memory exchanges of the tile done (NW warps): the last input boxes may be parked (kTmemX)
    //            +128 gate row of the tile the compute warps work on is in the table (helper; kHelperGate)
    // ring (the staging area): kTmemSlots slots of one 256-row TMA box each, used for loads while parking and for
    // stores while draining
    [[maybe_unused]] uint32_t tmem_base = 0;
    if constexpr (TMEM_IO) {
        constexpr int TBOXR = tmem_box_rows(NCOL);                         // rows per TMA box / ring slot
        constexpr uint32_t SLOTB = (uint32_t)TBOXR * ROWB;                 // bytes of one ring slot
        const uint32_t bar_in_full = bar, bar_in_free = bar + 8, bar_out_full = bar + 16, bar_out_free = bar + 24,
                       bar_landed = bar + 32;
        uint32_t *tmem_base_s = reinterpret_cast<uint32_t *>(smem_raw + bar_off + 96);
        if (tid == NT) {
            mbar_init(bar_in_full, 1);
            mbar_init(bar_in_free, NW);
            mbar_init(bar_out_full, NW);
            mbar_init(bar_out_free, 1);
            mbar_init(bar + 112, NW);                                  // inverse stage-0 read done (split barrier, sched bit 1)
            mbar_init(bar + 120, NW);                                  // both tensor-memory exchanges of the tile done (kTmemX)
            mbar_init(bar + 128, 1);                                   // gate row of the tile in work staged (helper; kHelperGate)
#pragma unroll
            for (int i = 0; i < kTmemSlots; ++i) mbar_init(bar_landed + 8 * i, 1);
        }
        if (tid >= NT && tid < NT + 32) tmem_alloc(smem_u32(tmem_base_s), 512);   // helper warp 0 owns the allocation
        tc_fence_before();
        __syncthreads();
        tc_fence_after();
        tmem_base = *tmem_base_s;
        griddep_wait();
        griddep_launch();
        if (tid >= NT) {
            asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kTmemHelperRegs));
            const int hl = tid - NT;                              // 0..127 = TMEM lane this helper thread serves
            const uint32_t tq = tmem_base + ((uint32_t)(hl & ~31) << 16);
            constexpr int NBOX = N / TBOXR;                       // TMA boxes per tile
            constexpr int GPB = TBOXR * NCOL / 128;               // 128-element groups per box
            [[maybe_unused]] constexpr int kGateStep = NBOX >= 3 ? 2 : NBOX - 1;   // kHelperGate: step after which the staged gate row is published
#ifndef SPX_TMEM_SP
#define SPX_TMEM_SP 1
#endif
            constexpr int SP = SPX_TMEM_SP;                       // TMA stores allowed to be still reading their slot
            constexpr int DEPTH = kTmemSlots - SP;                // loads kept in flight
            const bool idle = (p.sched & 8) != 0;                 // diagnostic: no tile I/O at all (results invalid)
            // The helper works in phases.  Phase P parks tile P of this CTA in TMEM-IN and drains tile P - 2 from TMEM-OUT,
            // one TMA box each per step, through ONE ring: a slot receives a load, is read out into tensor memory, is refilled
            // by the same threads with a box of results, leaves through a TMA store and then takes the next load.  Loads
            // therefore stay in flight while results drain (DEPTH boxes ahead), and the load stream runs across tile
            // boundaries.  Load box idx (stream index over all tiles of this CTA) is consumed at step idx, in slot idx % slots.
            // tile sequence of this CTA: tile_of(s), s = 0, 1, ...; paired: tiles 2P, 2P+1 of pair P = blockIdx.x + (s / 2) gridDim.x
            const bool pair = !DGATE && p.pair_tiles != 0;
            auto tile_of = [&](int s_) {
                if constexpr (DGATE) return (int)blockIdx.x + (s_ >> 1) * (int)gridDim.x;      // every tile twice: V rows, then dY rows
                return pair ? 2 * ((int)blockIdx.x + (s_ >> 1) * (int)gridDim.x) + (s_ & 1) : (int)blockIdx.x + s_ * (int)gridDim.x;
            };
            const int units = pair ? p.num_tiles / 2 : p.num_tiles;     // pairs or tiles dealt round robin to the CTAs
            const int my_units = ((int)blockIdx.x < units && !idle) ? (units - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
            const int my_tiles = (pair || DGATE) ? 2 * my_units : my_units;
            const int total_boxes = my_tiles * NBOX;
            // load stream state (used by the elected thread): next box to issue, kept incrementally -- the elected thread's
            // path between two barriers is the helper's critical path, so no divisions there except once per tile
            int ld_left = total_boxes, ld_k = 0, ld_t = tile_of(0), ld_tb = 0, ld_tc = 0, ld_seq = 0;
            uint32_t ld_slot = 0;
            constexpr int TCH = DIT ? CH : NCOL * CH;             // channels per tile
            if (my_tiles > 0) {
                ld_tb = ld_t / p.tiles_per_row;
                ld_tc = (ld_t - ld_tb * p.tiles_per_row) * TCH;
            }
            auto issue_next = [&]() {                             // load thread: next box of the stream -> next slot
                // (no proxy fence: the slot was last touched by shared-memory READS of the helper threads, ordered by the
                // helper barrier, or by a TMA store whose read the store thread has waited for)
                mbar_expect_tx(bar_landed + 8 * ld_slot, SLOTB);
                if constexpr (DIT)
                    tma_load_4d(smem_u32(stg) + ld_slot * SLOTB, &tmap, ld_tc, 0, ld_k * TBOXR, ld_tb, bar_landed + 8 * ld_slot);
                else
                    tma_load_3d(smem_u32(stg) + ld_slot * SLOTB, (DGATE && (ld_seq & 1)) ? &tmap_out : &tmap, ld_tc, ld_k * TBOXR, ld_tb,
                                bar_landed + 8 * ld_slot);
                --ld_left;
                ld_slot = (ld_slot + 1 == kTmemSlots) ? 0 : ld_slot + 1;
                if (++ld_k == NBOX) {
                    ld_k = 0;
                    ld_t = tile_of(++ld_seq);
                    ld_tb = ld_t / p.tiles_per_row;
                    ld_tc = (ld_t - ld_tb * p.tiles_per_row) * TCH;
                }
            };
            // two duty threads in different warps so the store and the load of a step are issued side by side: kStoreLane
            // issues / commits / waits for the TMA stores, kLoadLane issues the loads.  The load of step j goes to the slot of the
            // store of step j - 2, whose read the store thread awaited before it entered the barrier of step j.
#if SPX_SPLIT_DUTY
            constexpr int kStoreLane = 0, kLoadLane = 32, kDepth = kTmemSlots - 2;
#else
            constexpr int kStoreLane = 0, kLoadLane = 0, kDepth = DEPTH;
#endif
            if (hl == kLoadLane) {
                for (int i = 0; i < kDepth && ld_left > 0; ++i) issue_next();
                if (p.prefetch == 1 && my_tiles > 1) prefetch_tile(tile_of(1));
            }
            // prefetch == 2: the four CTAs that work on four adjacent channel tiles pull the NEXT tiles' rows into L2 as whole
            // 128-byte lines (one quarter of the rows each), so DRAM sees full-line reads instead of 32-byte pieces
#if SPX_COOP_PF
            const bool coop_pf = (p.prefetch == 2) && (p.tiles_per_row % 4 == 0) && (gridDim.x % 4 == 0) && (NCOL * CH == 8) &&
                                 (sizeof(TIN) == 4) && (p.n_in == N) && hl < 64;
#endif
            uint32_t sl = 0, landed_par = 0;                      // ring slot of the current step; one landed-parity bit per slot
            const int phases = my_tiles > 0 ? my_tiles + (DGATE ? 0 : 2) : 0;
#pragma unroll 1
            for (int P = 0; P < phases; ++P) {
                const bool do_park = P < my_tiles, do_drain = !DGATE && P >= 2;
#if SPX_HELPER_TL
                // diagnostic build: the elected thread records, per phase, entry / start / end stamps (ns) and the cycles it
                // spent waiting for landings, moving data, at the helper barrier and in its TMA duties
                const bool tl_on = p.timeline && hl == 0 && P < kTimelineTiles;
                unsigned long long *tl = p.timeline + (((size_t)blockIdx.x * kTimelineGroups + 4) * kTimelineTiles + (P < kTimelineTiles ? P : 0)) * kTimelineSlots;
                long long c_land = 0, c_move = 0, c_bar = 0, c_tma = 0;
                if (tl_on) tl[0] = globaltimer_ns();
#endif
                // kHelperGate: while the compute warps work on tile P - 1 the helper stages that tile's gate row.  They have pulled the tile
                // (this wait), so they are past the barrier that ends inverse stage 1 of tile P - 2: nobody reads the table any more.  The
                // copies (LDGSTS, no registers) land under the first steps; after step kGateStep the helper rescales its own entries
                // and publishes the table on barrier +128, which the compute warps await before their middle pass (~ 3 us from here)
                const bool do_gate = kHelperGate && hgate && P >= 1 && P <= my_tiles;
                if ((do_park || do_gate) && P >= 1) mbar_wait(bar_in_free, (P - 1) & 1);    // compute warps pulled tile P-1 out of TMEM-IN
                if constexpr (kHelperGate) {
                    if (do_gate) {
                        const int tg = tile_of(P - 1);
                        const int gb = tg / p.tiles_per_row, gc = (tg - gb * p.tiles_per_row) * TCH;
                        const int gg = p.gw_shift >= 0 ? (gc >> p.gw_shift) : gc / p.group_width;
                        gate_copy_async_rolled<N>(gate_s, p.gate + ((long long)gb * p.NG + gg) * (N / 2 + 1), hl);
                    }
                }
                if (do_drain) mbar_wait(bar_out_full, (P - 2) & 1);            // results of tile P-2 sit in TMEM-OUT
                tc_fence_after();
                if (p.prefetch == 1 && hl == kLoadLane && P + 2 < my_tiles) prefetch_tile(tile_of(P + 2));
#if SPX_HELPER_TL
                if (tl_on) tl[1] = globaltimer_ns();
#endif
                const int td = do_drain ? tile_of(P - 2) : 0;
                const int tb = do_drain ? td / p.tiles_per_row : 0;
                const int tc = do_drain ? (td - tb * p.tiles_per_row) * TCH : 0;
                // cooperative prefetch target: rows of this CTA's tile P + 1 (its loads are issued one phase from now)
#if SPX_COOP_PF
                const TIN *pf = nullptr;
                if (coop_pf && P + 1 < my_tiles) {
                    const int tp = tile_of(P + 1);
                    const int pb = tp / p.tiles_per_row, pc = tp - pb * p.tiles_per_row;
                    pf = vbase + (long long)pb * p.v_sb + (long long)((pc & 3) * 64 + hl) * p.v_sn + (pc & ~3) * (NCOL * CH);
                }
#endif
#pragma unroll 1
                for (int k = 0; k < NBOX; ++k) {
                    // ring entries hold the tensor's own element type (4 fp32 or 4 bf16 channels); tensor memory always fp32
                    LT *slot = reinterpret_cast<LT *>(stg + sl * SLOTB);
                    Cx<float2> v[GPB];
#if SPX_COOP_PF
                    if (pf) {
                        prefetch_l2(pf);
                        pf += (long long)TBOXR * p.v_sn;
                    }
#endif
#if SPX_HELPER_TL
                    long long c0 = clock64(), c1 = c0;
#endif
                    if (do_park) {
                        if constexpr (kTmemX) {
                            // the cells of the last input boxes are the compute warps' exchange columns while they work on tile P - 1
                            if (k == NBOX - kTmemXBoxes && P >= 1) {
                                mbar_wait(bar + 120, (P - 1) & 1);
                                tc_fence_after();
                            }
                        }
                        mbar_wait(bar_landed + 8 * sl, (landed_par >> sl) & 1);
                        landed_par ^= 1u << sl;
#if SPX_HELPER_TL
                        c1 = clock64();
#endif
#pragma unroll
                        for (int g = 0; g < GPB; ++g) v[g] = Lin<MODE, TIN>::get(slot[g * 128 + hl]);   // consecutive lanes, consecutive entries
#pragma unroll
                        for (int g = 0; g < GPB; ++g) tmem_st4(tq + (uint32_t)(4 * (k * GPB + g)), v[g].re.x, v[g].re.y, v[g].im.x, v[g].im.y);
                    }
                    if (do_drain) {
#pragma unroll
                        for (int g = 0; g < GPB; ++g) tmem_ld4(tq + (uint32_t)(TCOLS + 4 * (k * GPB + g)), v[g].re.x, v[g].re.y, v[g].im.x, v[g].im.y);
                        tmem_wait_ld();
                        // every thread refills exactly the ring entries it has just read: no barrier in between
#pragma unroll
                        for (int g = 0; g < GPB; ++g) slot[g * 128 + hl] = Lin<MODE, TOUT>::put(v[g]);
                        fence_proxy_async();
                    }
                    if constexpr (kHelperGate) {
                        if (do_gate && k == kGateStep) {           // own copies have landed: 1/n_fft, imag(DC) = imag(Nyquist) = 0
                            cp_async_wait_all();
                            gate_scale_own_rolled<N>(gate_s, hl, p.inv_n);
                        }
                    }
#if SPX_LAZY_STORE_WAIT && SPX_SPLIT_DUTY
                    // the load issued after this barrier reuses the slot of the store of step - 2: that store's read must be over.
                    // Waiting here, behind this step's data movement, instead of right after the commit keeps the elected thread off
                    // the critical path (all but the latest SP committed stores have been read)
                    if (hl == kStoreLane) tma_wait_read<SP>();
#endif
#if SPX_HELPER_TL
                    const long long c2 = clock64();
#endif
                    helper_bar();                                  // slot read by all (park) / written by all (drain)
#if SPX_HELPER_TL
                    const long long c3 = clock64();
#endif
                    if constexpr (kHelperGate) {
                        if (do_gate && k == kGateStep && hl == 0) mbar_arrive(bar + 128);
                    }
                    if (hl == kStoreLane && do_drain) {
                        if constexpr (DIT) tma_store_4d(&tmap_out, smem_u32(slot), tc, 0, k * TBOXR, tb);
                        else tma_store_3d(&tmap_out, smem_u32(slot), tc, k * TBOXR, tb);
                        tma_commit();
#if !(SPX_LAZY_STORE_WAIT && SPX_SPLIT_DUTY)
                        tma_wait_read<SP>();                       // the store of step - SP has left its slot
#endif
                    }
                    if (hl == kLoadLane && ld_left > 0) issue_next();   // into a slot whose store is known to have been read
#if SPX_HELPER_TL
                    c_land += c1 - c0; c_move += c2 - c1; c_bar += c3 - c2; c_tma += clock64() - c3;
#endif
                    sl = (sl + 1 == kTmemSlots) ? 0 : sl + 1;
                }
#if SPX_HELPER_TL
                if (tl_on) { tl[2] = globaltimer_ns(); tl[3] = c_land; tl[4] = c_move; tl[5] = c_bar; tl[6] = c_tma; }
#endif
                if (do_park) tmem_wait_st();
                tc_fence_before();
                helper_bar();
                if (hl == 0) {
                    if (do_park) mbar_arrive(bar_in_full);         // tile P is parked
                    if (do_drain) mbar_arrive(bar_out_free);       // TMEM-OUT may be overwritten (stores still draining smem)
                }
            }
            if (hl == 0) tma_wait_all();
        } else {
            asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kTmemComputeRegs));
        }
    }
    if constexpr (!TMEM_IO) {
        griddep_wait();
        griddep_launch();
    }
    if constexpr (TMA_IN && !TMEM_IO) {
        if (tid == (SEP ? NT : 0)) {
            mbar_init(bar, 1);
            mbar_init(bar_buf_free, NW);
            mbar_init(bar_stg_full, NW);
            mbar_init(bar_stg_full + 8, NW);
            mbar_init(bar_stg_free, 1);
            mbar_init(bar_stg_free + 8, 1);
            if ((int)blockIdx.x < p.num_tiles) issue_tile_load(blockIdx.x);
        }
        __syncthreads();   // all threads, once: barriers are initialised
        if constexpr (SEP) {
            if (tid >= NT) {
                asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
                // ---------------- TMA producer warpgroup: every bulk-tensor load / store of this CTA is issued here, so
                // no compute warp ever waits for the TMA queue.  One elected lane does the work.
                if (tid == NT) {
                    int it = 0;
                    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
                        producer_next_load(tile, it);
                        for (int j = 0; j < NR; ++j) producer_store_round(tile, j);
                        if (p.prefetch && tile + 2 * (int)gridDim.x < p.num_tiles) prefetch_tile(tile + 2 * gridDim.x);
                    }
                    tma_wait_all();   // staging buffers must outlive the last TMA stores
                }
                return;
            }
            asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(compute_regs(NT, MINB)));
        }
    }
    if (!TMEM_IO || tid < NT) {
    int tile_it = 0;
    uint32_t parity = 0;
    int tl_tile = 0;
#define SPX_MARK(slot)                                                                                   \
    if (p.timeline && (tid & 127) == 0 && tl_tile < kTimelineTiles)                                      \
        p.timeline[(((size_t)blockIdx.x * kTimelineGroups + (tid >> 7)) * kTimelineTiles + tl_tile) * kTimelineSlots + (slot)] = globaltimer_ns();
    uint32_t rnd = 0;   // output staging rounds issued so far (buffer = rnd & 1)
    // stage-0 butterfly of (thread, iteration): channel column and row offset u.  The TMEM variant ties u to the TMEM lane
    // the thread can reach: lane = 32 * (warp % 4) + lane id = u mod 128.
    auto map0 = [&](int it, int &col, int &u) {
        if constexpr (TMEM_IO) {
            // flat element index e = row * NCOL + col lives in TMEM lane e % 128: this thread's lane fixes (col, u mod 128/NCOL)
            const int wq = tid >> 5, tl = 32 * (wq & 3) + (tid & 31);
            col = tl % NCOL;
            u = tl / NCOL + (128 / NCOL) * (wq >> 2);
        } else {
            const int w = tid + it * NT;
            col = w % NCOL;
            u = w / NCOL;
        }
    };

    // tile -> (row of the [B'][n_fft][C] tensor, channel-tile column), advanced incrementally: integer divisions sit on every
    // warp's serial path between two tiles, so the loop has none (group index by shift when group_width is a power of two)
    // Paired order (TMEM variant, p.pair_tiles): tiles 2P and 2P+1 of pair P = blockIdx.x + k gridDim.x back to back -- the two
    // channel tiles of one gate group, so the gate row is staged (or, ANCH, generated) once per pair.
    const bool pair = TMEM_IO && !DGATE && p.pair_tiles != 0;
    const int stride = pair ? 2 * (int)gridDim.x : (int)gridDim.x;
    const int step_row = stride / p.tiles_per_row, step_col = stride - step_row * p.tiles_per_row;
    const int first_tile = pair ? 2 * (int)blockIdx.x : (int)blockIdx.x;
    int brow = first_tile / p.tiles_per_row, tcol = first_tile - brow * p.tiles_per_row;
    int nrow_ = 0, ncol_ = 0, tile_next_ = 0, seq = 0;
    auto gdiv = [&](int x) { return p.gw_shift >= 0 ? (x >> p.gw_shift) : x / p.group_width; };
    for (int tile = first_tile; tile < p.num_tiles; tile = tile_next_, brow = nrow_, tcol = ncol_, ++seq) {
        const bool partner_next = pair && !(seq & 1);   // the next tile is the second half of this pair: same row, same gate row
        if (DGATE && !(seq & 1)) {                      // dgate: the same tile again, with dY's rows
            nrow_ = brow;
            ncol_ = tcol;
            tile_next_ = tile;
        } else if (partner_next) {
            nrow_ = brow;
            ncol_ = tcol + 1;
            tile_next_ = tile + 1;
        } else {
            nrow_ = brow + step_row;
            ncol_ = tcol + step_col - (pair ? 1 : 0);
            tile_next_ = tile + stride - (pair ? 1 : 0);
            if (ncol_ >= p.tiles_per_row) { ncol_ -= p.tiles_per_row; ++nrow_; }
        }
        const int b = SUB ? (brow >> p.sub_shift) : brow;        // batch row (gate / memory)
        const int qsub = SUB ? (brow & (p.sub_R - 1)) : 0;       // which interleaved sub-transform
        const int ce0 = DIT ? tcol : tcol * NCOL;                // first element (4-channel group) of the tile; DIT: the only one
        const int c0 = ce0 * CH;
        const int g0 = gdiv(c0);
        const bool full_tile = (ce0 + NCOL <= CE) && (p.n_in == N);
        SPX_MARK(0)

        // ---- stage the gate tables of this tile (asynchronous copies, awaited before the barrier that follows stage 0).
        // With one table per tile this was already started while the previous tile finished (see below); else do it here.
        if constexpr (!RFFT_ONLY && !DGATE) {
            if constexpr (DIT) {
                if (seq == 0) {   // half spectrum of the 2 N-point transform: N + 1 entries in the two table slots
                    if constexpr (kDitAsync) {
                        gate_copy_async<NTOT, NT, GKD>(gate_s, p.gate + ((long long)b * p.NG + g0) * (NTOT / 2 + 1), tid);
                    } else {
                        float2 gv[GKD];
                        gate_fetch<NTOT, NT, GKD>(gv, gate_row<ANCH>(p, b, g0, NTOT / 2 + 1), tid);
                        gate_put<NTOT, NT, GKD>(gate_s, gv, tid, p.inv_n);
                    }
                }
            } else if constexpr (SUB) {
                if (seq == 0) {
                    float2 gv[GKS];
                    gate_fetch_sub<N, NT, GKS>(gv, gate_row<ANCH>(p, b, g0, (N * p.sub_R) / 2 + 1), tid, qsub, p.sub_R);
                    gate_put_sub<N, NT, GKS>(gate_s, gv, tid, qsub, p.sub_R, p.inv_n);
                }
            } else if (hgate) {
                // the helper warpgroup staged this tile's row before it announced the tile
            } else if (!gate_early || seq == 0) {
                for (int t = 0; t < p.gate_tables; ++t) {
                    const int g = g0 + t;
                    if (g < p.NG) {
                        if constexpr (kGateAsync) {
                            gate_copy_async<N, NT, GK>(gate_s + t * GS, p.gate + ((long long)b * p.NG + g) * (N / 2 + 1), tid);
                        } else {
                            float2 gv[GK];
                            gate_fetch<N, NT, GK>(gv, gate_row<ANCH>(p, b, g, N / 2 + 1), tid);
                            gate_put<N, NT, GK>(gate_s + t * GS, gv, tid, p.inv_n);
                        }
                    }
                }
            }
        }

        // ---- forward stage 0: (TMA landing buffer | global) -> butterfly -> twiddle -> padded column layout
        {
            Cx<V> x0[ITERS0][R0];
            if constexpr (TMEM_IO) {
                // the helper warpgroup parked this tile in TMEM-IN: row u + L0 m sits in lane u % 128, column group
                // ((u NCOL + col) / 128 + MSTRIDE / 4 * m); this warp's 32 lanes are exactly its 32 (column, u) pairs
                if (!(p.sched & 8)) mbar_wait(bar, tile_it & 1);
                tc_fence_after();
                SPX_MARK(1)
                int col, u;
                map0(0, col, u);
                const uint32_t ta = tmem_base + ((uint32_t)(32 * ((tid >> 5) & 3)) << 16) + (uint32_t)(4 * ((u * NCOL + col) >> 7));
#pragma unroll
                for (int m = 0; m < R0; ++m)
                    tmem_ld4(ta + (uint32_t)(m * MSTRIDE), x0[0][m].re.x, x0[0][m].re.y, x0[0][m].im.x, x0[0][m].im.y);
                tmem_wait_ld();
                tc_fence_before();
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(bar + 8);   // TMEM-IN consumed by this warp
                if (p.sched & 4) {
                    // diagnostic: tile I/O only -- hand the tile straight back to the helper warpgroup (results invalid)
                    if (tile_it >= 1) mbar_wait(bar + 24, (tile_it - 1) & 1);
                    tc_fence_after();
                    const uint32_t tb_ = ta + (uint32_t)TCOLS;
#pragma unroll
                    for (int m = 0; m < R0; ++m)
                        tmem_st4(tb_ + (uint32_t)(m * MSTRIDE), x0[0][m].re.x, x0[0][m].re.y, x0[0][m].im.x, x0[0][m].im.y);
                    tmem_wait_st();
                    tc_fence_before();
                    __syncwarp();
                    if ((tid & 31) == 0) mbar_arrive(bar + 16);
                    ++tile_it;
                    continue;
                }
            } else if constexpr (TMA_IN) {
                mbar_wait(bar, parity);
                parity ^= 1;
                SPX_MARK(1)
                const LT *lin = reinterpret_cast<const LT *>(smem_raw);
#pragma unroll
                for (int it = 0; it < ITERS0; ++it) {
                    const int w = tid + it * NT;
                    int col, u;
                    map0(it, col, u);
                    if (ITEMS0 % NT == 0 || w < ITEMS0) {
#pragma unroll
                        for (int m = 0; m < R0; ++m) x0[it][m] = Lin<MODE, TIN>::get(lin[(u + m * L0) * NCOL + col]);
                    }
                }
                cta_sync<NT, SEP>();   // every landed element is in registers: the buffer may be rewritten in place
            } else {
                const TIN *vb = vbase + (long long)brow * p.v_sb + c0;
#pragma unroll
                for (int it = 0; it < ITERS0; ++it) {
                    const int w = tid + it * NT;
                    int col, u;
                    map0(it, col, u);
                    if (ITEMS0 % NT == 0 || w < ITEMS0) {
                        const TIN *vp = vb + (long long)u * p.v_sn + col * CH;
                        if (full_tile) {
#pragma unroll
                            for (int m = 0; m < R0; ++m) x0[it][m] = GIO<MODE, TIN>::load(vp + (long long)(m * L0) * p.v_sn);
                        } else {
                            // ragged tile: read a clamped (always valid) address, then select; the select reads the
                            // loaded register, so the loads still overlap (a predicated load + zero fill would not)
                            const bool colok = (ce0 + col) < CE;
                            const TIN *vq = vb + (colok ? col * CH : 0);
#pragma unroll
                            for (int m = 0; m < R0; ++m) {
                                const int row = u + m * L0;
                                const bool ok = colok && row < p.n_in;
                                Cx<V> t = GIO<MODE, TIN>::load(vq + (long long)(ok ? row : 0) * p.v_sn);
                                if (!ok) t = {vzero(V()), vzero(V())};
                                x0[it][m] = t;
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int it = 0; it < ITERS0; ++it) {
                const int w = tid + it * NT;
                int col, u;
                map0(it, col, u);
                if (ITEMS0 % NT == 0 || w < ITEMS0) {
                    Dft<R0, V>::run(x0[it]);
                    apply_twiddles<PL, 0, V>(x0[it], tw, p.tw, u);
                    if constexpr (TMEM_IO && !DGATE) {
                        // split barrier: every warp has pulled the previous tile's last pass out of the buffer
                        if ((p.sched & 2) && tile_it >= 1) mbar_wait(bar + 112, (tile_it - 1) & 1);
                    }
                    S *cb = buf + col * CS + u + (u >> 4);
#pragma unroll
                    for (int q = 0; q < R0; ++q) cb[q * L0 + ((q * L0) >> 4)] = E::pack(x0[it][q]);
                }
            }
        }
        if constexpr (kDitAsync) cp_async_wait_all();   // this thread's share of the (raw) gate row has landed; the barrier publishes it
        if constexpr (kGateAsync) {
            // this thread's share of the gate tables has landed: rescale it in place before the barrier publishes it
            cp_async_wait_all();
            if constexpr (!kGateRaw) {
                for (int t = 0; t < p.gate_tables; ++t)
                    if (g0 + t < p.NG) gate_scale_own<N, NT, GK>(gate_s + t * GS, tid, p.inv_n);
            }
        }
        cta_sync<NT, SEP>();
        SPX_MARK(2)
        // stagger: with every warp in lock step all of them load, then all compute, then all store; holding back two of
        // the four warps of each scheduler lets their shared-memory phases overlap the others' butterflies
        if (TMEM_IO || (p.sched & 64)) stagger(p.skew_ns, tid);

        // ---- forward stages 1 .. NS-2: smem -> butterfly -> twiddle -> smem (in place)
        // The exchange between stage NS-2 and the middle pass is a 16x16 transpose among the 16 threads that share
        // (column, leading digits): with consecutive butterflies on consecutive lanes those threads are one half-warp,
        // in this pass and in the middle pass alike, so __syncwarp() orders it and warps run on unsynchronised.
        // kTmemX: stage 1 stays in registers (xk), is transposed through tensor memory and feeds the middle pass directly
        [[maybe_unused]] Cx<V> xk[kTmemX ? 16 : 1];
        [[maybe_unused]] uint32_t tmx = 0;           // this warp's 16 exchange columns: TMEM-IN cells of the last four input boxes
        [[maybe_unused]] S *cb1 = nullptr;           // stage-1 butterfly (column tid % 2, Q = warp, u = lane / 2): its 16 buffer slots
        if constexpr (kTmemX) {
            static_assert(PL::R(1) == 16 && PL::L(1) == 16 && RL == 16 && NCOL == 2, "tensor-memory exchange: 16 x 16 x 16 plan, two element columns");
            const int col1 = tid & 1, bf1 = tid >> 1, e1 = (bf1 >> 4) * 256 + (bf1 & 15);
            cb1 = buf + col1 * CS + e1 + (e1 >> 4);
            tmx = tmem_base + ((uint32_t)(32 * ((tid >> 5) & 3)) << 16) + (uint32_t)(TCOLS - 16 * kTmemXBoxes + 16 * (tid >> 7));
#pragma unroll
            for (int m = 0; m < 16; ++m) xk[m] = E::unpack(cb1[m * 16 + m]);
            Dft<16, V>::run(xk);
            apply_twiddles<PL, 1, V>(xk, tw, p.tw, bf1 & 15);
            tmem_xchg_fwd(xk, tmx);                  // (every warp has pulled its tile out of TMEM-IN: the barrier after stage 0)
            if constexpr (!kTmemXI) {
                tc_fence_before();
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(bar + 120);   // the helper may park the next tile's last boxes over the exchange columns
            }
        } else
        if constexpr (NS > 2) { fwd_inner_pass<PL, MODE, NCOL, NT, 1, kIlv>(buf, tw, p.tw, tid); if constexpr (NS == 3 && (kWarpLocal || kWarpLocalN)) __syncwarp(); else cta_sync<NT, SEP>(); }
        if constexpr (NS > 3) { fwd_inner_pass<PL, MODE, NCOL, NT, 2, kIlv>(buf, tw, p.tw, tid); if constexpr (kWarpLocal) __syncwarp(); else cta_sync<NT, SEP>(); }

        SPX_MARK(3)
        if constexpr (kHelperGate) {
            if (hgate) mbar_wait(bar + 128, tile_it & 1);   // the helper warpgroup has staged this tile's gate row
        }
        // ---- middle pass: last forward butterfly -> gate (+memory) -> first inverse butterfly
        {
            constexpr int NBF = N / RL, ITEMS = NCOL * NBF;
            constexpr int IPT = kWarpLocalN ? 16 / RL : 1;      // narrow-radix mapping: middle-pass items per thread
            for (int w = tid, jj = 0; kWarpLocalN ? (jj < IPT) : (w < ITEMS); w += NT, ++jj) {
                int col, Q;
                if constexpr (kWarpLocalN) {
                    // this thread's stage-1 butterfly is (col, Qf, u); the group's 16 items are 16 Qf + q, thread u takes
                    // q = IPT u + jj: consecutive lanes 16 elements apart (conflict-free with the 16-byte pad per 16 elements)
                    constexpr int NB1 = N / 16;
                    col = tid / NB1;
                    const int bf1 = tid - col * NB1, Qf = bf1 / RL, u1 = bf1 - Qf * RL;
                    Q = 16 * Qf + IPT * u1 + jj;
                } else if constexpr (kTmemX) {
                    // after the exchange: lane = 16 col + q1, register = u of stage 1; the warp is the leading digit
                    col = (tid >> 4) & 1;
                    Q = 16 * (tid >> 5) + (tid & 15);
                } else {
                    col = kIlv ? w % NCOL : w / NBF;
                    Q = kIlv ? w / NCOL : w - col * NBF;
                }
                if ((DIT ? ce0 : ce0 + col) >= CE) continue;   // column past the last channel: nothing to transform
                const int e0 = Q * RL;
                [[maybe_unused]] S *cb = buf + col * CS + e0 + (e0 >> 4);
                [[maybe_unused]] Cx<V> xl[kTmemX ? 1 : RL];
                Cx<V> (&x)[RL] = pick_ref<kTmemX>(xk, xl);
                if constexpr (!kTmemX) {
#pragma unroll
                    for (int m = 0; m < RL; ++m) x[m] = E::unpack(cb[m]);
                }
                Dft<RL, V>::run(x);
                const int klow = klow_of<PL>(Q);
                const int cabs = (DIT ? ce0 : ce0 + col) * CH;     // first channel of this element
                if constexpr (RFFT_ONLY) {
                    // half spectrum out (spectre.py:506 / :777): bins k <= n_fft/2 of this channel
                    float2 *sp = reinterpret_cast<float2 *>(p.out) + (long long)b * p.o_sb + cabs;
#pragma unroll
                    for (int q = 0; q < RL; ++q) {
                        const int k = klow + PLAST * q;
                        if (k <= N / 2) sp[(long long)k * p.o_sn] = make_float2(x[q].re, x[q].im);
                    }
                    continue;
                }
                if constexpr (DGATE) {
                    // this thread's private stash: its own tensor-memory lane, 64 columns of the (otherwise unused) TMEM-OUT half
                    const uint32_t ts = tmem_base + ((uint32_t)(32 * ((tid >> 5) & 3)) << 16) + (uint32_t)(TCOLS + 64 * (tid >> 7));
                    if (!(seq & 1)) {
                        // V's rows: keep the packed spectrum Z[k], k = klow + PLAST q, until dY's spectrum of the same tile is here
#pragma unroll
                        for (int q = 0; q < RL; ++q) tmem_st4(ts + 4 * q, x[q].re.x, x[q].re.y, x[q].im.x, x[q].im.y);
                        tmem_wait_st();
                    } else {
                        // dY's rows: T[k] = sum over the two packed lanes of conj(Z[k]) W[k]  (Z = v_a + i v_b, W = d_a + i d_b; the
                        // per-channel products conj(V_c) D_c are recovered below from T[k] + conj(T[N - k]))
                        float2 t[RL];
#pragma unroll
                        for (int q = 0; q < RL; ++q) {
                            Cx<V> z;
                            tmem_ld4(ts + 4 * q, z.re.x, z.re.y, z.im.x, z.im.y);
                            tmem_wait_ld();
                            const float2 re2 = __ffma2_rn(z.im, x[q].im, __fmul2_rn(z.re, x[q].re));                     // Zre Wre + Zim Wim
                            const float2 im2 = __ffma2_rn(make_float2(-z.im.x, -z.im.y), x[q].re, __fmul2_rn(z.re, x[q].im));   // Zre Wim - Zim Wre
                            t[q] = make_float2(re2.x + re2.y, im2.x + im2.y);
                        }
                        // (one item per thread and full tiles in this variant -- the host checks C % tile channels == 0 -- so every
                        // thread of the CTA reaches this barrier exactly once)
                        cta_sync<NT, SEP>();                 // every warp has pulled its middle-pass inputs out of the buffer
                        // T in natural bin order, padded like the gate table (k + (k >> 4)): conflict-free to write (lanes 16 bins
                        // apart) and to read (consecutive bins)
                        float2 *tn = reinterpret_cast<float2 *>(buf) + (size_t)col * (N + (N >> 4) + 1);
#pragma unroll
                        for (int q = 0; q < RL; ++q) {
                            const int k = klow + PLAST * q;
                            tn[k + (k >> 4)] = t[q];
                        }
                    }
                    continue;
                }
                if constexpr (DIT) {
                    // x[q] = S_c[k'], k' = klow + PLAST q: the N-point spectrum of the rows 2 n + c (c = col = lane parity).  The two
                    // lanes of a pair trade halves so that lane c owns q in [8 c, 8 c + 8) of BOTH sub-spectra (32 shuffles each
                    // way, same work on both lanes), then for each of its bins:
                    //   X[k'] = S0 + W^k' S1,  X[k' + N] = S0 - W^k' S1          (W = exp(-2 pi i / 2N): decimation in time)
                    //   Y = Gfull X (+ memory)                                   (Gfull[k' + N] = conj(G[N - k']), Hermitian)
                    //   S0' = Y[k'] + Y[k' + N],  S1' = conj(W^k') (Y[k'] - Y[k' + N])
                    static_assert(RL == 16, "DIT2 middle pass: radix-16 last stage");
                    const bool odd = (col & 1) != 0;
                    // component-wise register selects (a struct-valued ?: on array elements becomes a pointer select and demotes the
                    // array to local memory)
                    auto csel = [](bool c, const Cx<V> a, const Cx<V> b) {
                        Cx<V> r;
                        r.re.x = c ? a.re.x : b.re.x; r.re.y = c ? a.re.y : b.re.y;
                        r.im.x = c ? a.im.x : b.im.x; r.im.y = c ? a.im.y : b.im.y;
                        return r;
                    };
                    auto shx = [](const Cx<V> &a) {
                        Cx<V> r;
                        r.re.x = __shfl_xor_sync(0xffffffffu, a.re.x, 1); r.re.y = __shfl_xor_sync(0xffffffffu, a.re.y, 1);
                        r.im.x = __shfl_xor_sync(0xffffffffu, a.im.x, 1); r.im.y = __shfl_xor_sync(0xffffffffu, a.im.y, 1);
                        return r;
                    };
                    float wbs, wbc;                                    // W^klow = exp(-i pi klow / N)
                    sincospif(-(float)klow * (1.0f / (float)N), &wbs, &wbc);
                    const float2 *gs = gate_s;
                    // one bin pair at a time (trade, combine, gate, split, trade back): only 16 + ~6 packed values are ever live
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const Cx<V> recv = shx(csel(odd, x[j], x[j + 8]));
                        const Cx<V> s0 = csel(odd, recv, x[j]);
                        const Cx<V> s1 = csel(odd, x[j + 8], recv);
                        // W^(PLAST q) = W_32^q, q = j + 8 odd: W_32^(j + 8) = -i W_32^j
                        constexpr float c32[8] = {1.f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                                                  0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f,
                                                  0.19509032201612826785f};
                        constexpr float s32[8] = {0.f, -0.19509032201612826785f, -0.38268343236508977173f, -0.55557023301960222474f,
                                                  -0.70710678118654752440f, -0.83146961230254523708f, -0.92387953251128675613f,
                                                  -0.98078528040323044913f};
                        float wr = wbc * c32[j] - wbs * s32[j], wi = wbc * s32[j] + wbs * c32[j];
                        if (odd) { const float t_ = wr; wr = wi; wi = -t_; }
                        const int kp = klow + PLAST * (j + (odd ? 8 : 0));       // k' in [0, N)
                        const int km = N - kp;                                   // mirror partner of bin k' + N, in [1, N]
                        const Cx<V> t = cmul(s1, wr, wi);
                        Cx<V> lo = cadd(s0, t), hi = csub(s0, t);
                        float2 glo = gs[kp + (kp >> 4)], ghi = gs[km + (km >> 4)];
                        if constexpr (kDitAsync) {   // raw row in the table: 1/n here; imag of DC (kp = 0) and of Nyquist (km = N) ignored
                            glo.x *= p.inv_n; ghi.x *= p.inv_n;
                            glo.y = (kp == 0) ? 0.f : glo.y * p.inv_n;
                            ghi.y = (kp == 0) ? 0.f : ghi.y * p.inv_n;
                        }
                        lo = cmul(lo, glo.x, glo.y);
                        hi = cmulc(hi, ghi.x, ghi.y);
                        if (HAS_MEM) {
                            const float sg = (kp == 0) ? 0.f : p.inv_n;          // imag of DC (lo) and of Nyquist (hi) ignored
                            mem_add(lo, p.mem + (long long)kp * p.mem_stride + cabs, p.inv_n, sg, MODE);
                            mem_add(hi, p.mem + (long long)km * p.mem_stride + cabs, p.inv_n, -sg, MODE);
                        }
                        const Cx<V> r0 = cadd(lo, hi);                           // S0'[q]
                        const Cx<V> r1 = cmulc(csub(lo, hi), wr, wi);            // S1'[q]
                        const Cx<V> back = shx(csel(odd, r0, r1));
                        x[j] = cswap(csel(odd, back, r0));
                        x[j + 8] = cswap(csel(odd, r1, back));
                    }
                    Dft<RL, V>::run(x);
#pragma unroll
                    for (int m = 0; m < RL; ++m) cb[m] = E::pack(cswap(x[m]));
                    continue;
                }
                const float2 *gs = gate_s + (p.gate_tables == 1 ? 0 : (gdiv(cabs) - g0) * GS);
                // bins k = klow + PLAST q (q < RL/2) use G[k]; the mirrored half uses conj(G[N - k])
                const int plo = klow + (klow >> 4);                // padded index of bin klow
                const int nk = N - klow;                           // N - klow - PLAST q stays >= N/2 - ... >= 1
#pragma unroll
                for (int q = 0; q < RL; ++q) {
                    if constexpr (SUB) {
                        // bin k' = klow + PLAST q of this sub-spectrum = bin qsub + R k' of the long transform; the table
                        // already holds the Hermitian-extended, conjugated-where-needed gate
                        const int kp = klow + PLAST * q;
                        const float2 g = gs[(PLAST % 16 == 0) ? plo + PLAST * q + ((PLAST * q) >> 4) : kp + (kp >> 4)];
                        x[q] = cmul(x[q], g.x, g.y);
                        if (HAS_MEM) {
                            const int nt = N * p.sub_R, kf = qsub + p.sub_R * kp;
                            const bool lowf = kf <= nt / 2;
                            const int kk = lowf ? kf : nt - kf;
                            const float sgn = (kf == 0 || kf == nt / 2) ? 0.f : (lowf ? p.inv_n : -p.inv_n);
                            mem_add(x[q], p.mem + (long long)kk * p.mem_stride + cabs, p.inv_n, sgn, MODE);
                        }
                        x[q] = cswap(x[q]);
                        continue;
                    }
                    const bool lower = q < RL / 2;
                    int k, kp;
                    if (lower) {
                        k = klow + PLAST * q;
                        kp = (PLAST % 16 == 0) ? plo + PLAST * q + ((PLAST * q) >> 4) : k + (k >> 4);
                    } else {
                        k = nk - PLAST * q;
                        kp = k + (k >> 4);
                    }
                    float2 g = gs[kp];
                    if constexpr (kGateRaw) {
                        g.x *= p.inv_n;
                        g.y = ((q == 0 || q == RL / 2) && klow == 0) ? 0.f : g.y * p.inv_n;
                    }
                    x[q] = lower ? cmul(x[q], g.x, g.y) : cmulc(x[q], g.x, g.y);
                    if (HAS_MEM) {
                        const float sgn = ((q == 0 || q == RL / 2) && klow == 0) ? 0.f : (lower ? p.inv_n : -p.inv_n);
                        mem_add(x[q], p.mem + (long long)k * p.mem_stride + cabs, p.inv_n, sgn, MODE);
                    }
                    x[q] = cswap(x[q]);
                }
                Dft<RL, V>::run(x);
                if constexpr (!kTmemXI) {
#pragma unroll
                    for (int m = 0; m < RL; ++m) cb[m] = E::pack(cswap(x[m]));
                }
            }
        }
        if constexpr (DGATE) {
            if (seq & 1) {
                cta_sync<NT, SEP>();                         // T of both element columns is in the buffer
                // dgate[b, g, k] += w_k / (2 n) * sum over columns of (T[k] + conj(T[(N - k) mod N])), k <= N/2, fixed order;
                // imag(DC) = imag(Nyquist) = 0 (irfft ignores them, so their gradient is zero)
                const float2 *tn = reinterpret_cast<const float2 *>(buf);
                float2 *dg = reinterpret_cast<float2 *>(p.out) + ((long long)b * p.NG + g0) * (N / 2 + 1);
                constexpr int TS = N + (N >> 4) + 1;
                for (int k = tid; k <= N / 2; k += NT) {
                    const int km = (N - k) & (N - 1);
                    float sr = 0.f, si = 0.f;
#pragma unroll
                    for (int c = 0; c < NCOL; ++c) {
                        if ((ce0 + c) >= CE) continue;
                        const float2 a = tn[(size_t)c * TS + k + (k >> 4)], m = tn[(size_t)c * TS + km + (km >> 4)];
                        sr += a.x + m.x;
                        si += a.y - m.y;
                    }
                    const bool edge = (k == 0) || (k == N / 2);
                    const float sc = (edge ? 0.5f : 1.0f) * p.inv_n;
                    atomicAdd(&dg[k].x, sr * sc);
                    if (!edge) atomicAdd(&dg[k].y, si * sc);
                }
            }
            cta_sync<NT, SEP>();                             // the next visit's stage 0 rewrites the buffer
            ++tl_tile;
            ++tile_it;
            continue;
        }
        if constexpr ((kWarpLocal || kWarpLocalN) && !RFFT_ONLY) __syncwarp(); else cta_sync<NT, SEP>();
        SPX_MARK(4)
        if constexpr (RFFT_ONLY) continue;

        if constexpr (kTmemXI) {
            // back to the stage-1 layout; the middle pass left (im, re)-swapped data, exactly what inverse stage 1 works on.
            // BEFORE the gate prefetch below: tcgen05.wait::st is a fence that also waits for the thread's global loads in flight
            tmem_xchg_inv(xk, tmx);
            tc_fence_before();
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(bar + 120);   // the helper may park the next tile's last boxes over the exchange columns
        }
        // the gate table is free again: start fetching the next tile's gate row, park it in registers
        // across the inner inverse passes, and publish it before the last pass
        float2 gnext[SUB ? GKS : (DIT ? GKD1 : (kGateAsync ? 1 : GK))];
        [[maybe_unused]] float2 gnext2[kEarlyGT == 2 ? GK : 1];   // second table of a wide tile
#ifdef SPX_DIAG_NOGATE   // timing experiment only (wrong results): the gate table is never refreshed
        const bool fetch_next = false;
#else
        const bool fetch_next = gate_early && !hgate && tile_next_ < p.num_tiles && !partner_next;   // the partner tile reuses the table
#endif
        int nq = 0;
        if (fetch_next) {
            const int nrow = nrow_;
            const int ng = gdiv(ncol_ * (DIT ? CH : NCOL * CH));
            if constexpr (kDitAsync) {
                // nothing to park: the row is copied straight into the table after the barrier below
            } else if constexpr (DIT) {
                gate_fetch<NTOT, NT, GKD1>(gnext, gate_row<ANCH>(p, nrow, ng, NTOT / 2 + 1), tid);
            } else if constexpr (SUB) {
                nq = nrow & (p.sub_R - 1);
                gate_fetch_sub<N, NT, GKS>(gnext, gate_row<ANCH>(p, nrow >> p.sub_shift, ng, (N * p.sub_R) / 2 + 1), tid, nq, p.sub_R);
            } else {
                if constexpr (!kGateAsync) {
                    gate_fetch<N, NT, GK>(gnext, gate_row<ANCH>(p, nrow, ng, N / 2 + 1), tid);
                    if constexpr (kEarlyGT == 2) {
                        if (p.gate_tables == 2 && ng + 1 < p.NG) gate_fetch<N, NT, GK>(gnext2, gate_row<ANCH>(p, nrow, ng + 1, N / 2 + 1), tid);
                    }
                }
            }
        }

        // ---- inverse stages NS-2 .. 1: smem -> twiddle -> butterfly -> smem (in place)
        if constexpr (NS > 3) { inv_inner_pass<PL, MODE, NCOL, NT, 2, kIlv>(buf, tw, p.tw, tid); cta_sync<NT, SEP>(); }
        if constexpr (kTmemXI) {
            apply_twiddles<PL, 1, V>(xk, tw, p.tw, (tid >> 1) & 15);
            Dft<16, V>::run(xk);
#pragma unroll
            for (int m = 0; m < 16; ++m) cb1[m * 16 + m] = E::pack(cswap(xk[m]));
            cta_sync<NT, SEP>();
        } else
        if constexpr (NS > 2) { inv_inner_pass<PL, MODE, NCOL, NT, 1, kIlv>(buf, tw, p.tw, tid); cta_sync<NT, SEP>(); }
        // DIT2: the second part of the next tile's gate row is fetched here and parked across the last inverse pass (its registers are
        // free now: nothing else is live between the barrier above and that pass's loads), published after its butterflies
        [[maybe_unused]] float2 g2[DIT ? GKD2 : 1];
        if (fetch_next) {
            if constexpr (kDitAsync) {
                gate_copy_async<NTOT, NT, GKD>(gate_s, p.gate + ((long long)nrow_ * p.NG + gdiv(ncol_ * CH)) * (NTOT / 2 + 1), tid);
            } else if constexpr (DIT) {
                gate_put<NTOT, NT, GKD1>(gate_s, gnext, tid, p.inv_n);
                gate_fetch<NTOT, NT, GKD2, ANCH, GKD1>(g2, gate_row<ANCH>(p, nrow_, gdiv(ncol_ * CH), NTOT / 2 + 1), tid);
            } else if constexpr (SUB) {
                gate_put_sub<N, NT, GKS>(gate_s, gnext, tid, nq, p.sub_R, p.inv_n);
            } else {
                if constexpr (kGateAsync) {
                    // every warp is past the middle pass: the next tile's gate row may stream into the table (kEarlyGT == 1 here:
                    // kGateAsync kernels have 8-channel tiles, one table)
                    const int nrow = nrow_;
                    const int ng = gdiv(ncol_ * NCOL * CH);
                    gate_copy_async<N, NT, GK>(gate_s, p.gate + ((long long)nrow * p.NG + ng) * (N / 2 + 1), tid);
                } else {
                    gate_put<N, NT, GK>(gate_s, gnext, tid, p.inv_n);
                    if constexpr (kEarlyGT == 2) {
                        if (p.gate_tables == 2 && gdiv(ncol_ * NCOL * CH) + 1 < p.NG) gate_put<N, NT, GK>(gate_s + GS, gnext2, tid, p.inv_n);
                    }
                }
            }
        }
        SPX_MARK(5)
        if ((TMEM_IO || (p.sched & 64)) && (p.sched & 1)) {
            // sched bits 8..11: step of this second stagger in quarters of the first one's (0 = same step)
            const int q4 = (p.sched >> 8) & 15;
            stagger((q4 && p.skew_ns < 0) ? -(((-p.skew_ns) / 100000) * 100000 + (((-p.skew_ns) % 100000) * q4) / 4) : p.skew_ns, tid);
        }

        // ---- inverse stage 0: smem -> twiddle -> butterfly -> global
        {
            Cx<V> x0[ITERS0][R0];
#pragma unroll
            for (int it = 0; it < ITERS0; ++it) {
                const int w = tid + it * NT;
                int col, u;
                map0(it, col, u);
                if (ITEMS0 % NT == 0 || w < ITEMS0) {
                    const S *cb = buf + col * CS + u + (u >> 4);
#pragma unroll
                    for (int q = 0; q < R0; ++q) x0[it][q] = cswap(E::unpack(cb[q * L0 + ((q * L0) >> 4)]));
                }
            }
            if constexpr (TMEM_IO) {
                // the next tile's stage 0 writes this buffer as soon as a warp gets there: everyone must have read first
                if (p.sched & 2) {
                    __syncwarp();
                    if ((tid & 31) == 0) mbar_arrive(bar + 112);
                } else {
                    cta_sync<NT, SEP>();
                }
            } else if constexpr (TMA_IN) {
                // this warp's share of the tile now lives in registers; when all warps have said so the producer
                // hands the buffer to the TMA unit for the next tile
                fence_proxy_async();
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(bar_buf_free);
                if constexpr (!SEP) {
                    if (tid == 0) producer_next_load(tile, tile_it);
                }
            }
            TOUT *ob = obase + (long long)brow * p.o_sb + c0;
#pragma unroll
            for (int it = 0; it < ITERS0; ++it) {
                const int w = tid + it * NT;
                int col, u;
                map0(it, col, u);
                (void)col;
                if (ITEMS0 % NT == 0 || w < ITEMS0) {
                    apply_twiddles<PL, 0, V>(x0[it], tw, p.tw, u);
                    Dft<R0, V>::run(x0[it]);
                }
            }
            if constexpr (DIT) {
                // (every warp is past the middle pass since the barrier that ended inverse stage 1; the next one reads the table
                // after the barrier that ends the next tile's stage 0)
                if constexpr (!kDitAsync) {
                    if (fetch_next) gate_put<NTOT, NT, GKD2, GKD1>(gate_s, g2, tid, p.inv_n);
                }
            }
            SPX_MARK(6)
            if constexpr (TMEM_IO) {
                // park the results in TMEM-OUT (same lane / column-group geometry as the input side); the helper
                // warpgroup drains them to HBM while this CTA already transforms the next tile
                if (tile_it >= 1 && !(p.sched & 8)) mbar_wait(bar + 24, (tile_it - 1) & 1);   // previous tile's results have been drained
                tc_fence_after();
                int col, u;
                map0(0, col, u);
                const uint32_t ta = tmem_base + ((uint32_t)(32 * ((tid >> 5) & 3)) << 16) +
                                    (uint32_t)(TCOLS + 4 * ((u * NCOL + col) >> 7));
#pragma unroll
                for (int m = 0; m < R0; ++m) {
                    const Cx<V> y = cswap(x0[0][m]);
                    tmem_st4(ta + (uint32_t)(m * MSTRIDE), y.re.x, y.re.y, y.im.x, y.im.y);
                }
                tmem_wait_st();
                tc_fence_before();
                __syncwarp();
                if ((tid & 31) == 0) mbar_arrive(bar + 16);   // results parked
            } else if constexpr (TMA_IN) {
                // results leave through two shared staging buffers (STG_MB row blocks each) and TMA stores; a buffer is
                // refilled only after the TMA unit has read it (mbarrier `free`), so stores of one round overlap the
                // fill of the next and the landing of the next tile
                using OT = typename Lin<MODE, TOUT>::T;
                constexpr int MB = SM::STG_MB;
                static_assert(R0 % MB == 0, "row blocks per staging round must divide the stage-0 radix");
#pragma unroll
                for (int j = 0; j < NR; ++j) {
                    const uint32_t kbuf = rnd & 1;
                    if (rnd >= 2) mbar_wait(bar_stg_free + 8 * kbuf, ((rnd >> 1) - 1) & 1);
                    OT *sb = reinterpret_cast<OT *>(stg + kbuf * STGB);
#pragma unroll
                    for (int it = 0; it < ITERS0; ++it) {
                        const int w = tid + it * NT;
                        int col, u;
                        map0(it, col, u);
                        if (ITEMS0 % NT == 0 || w < ITEMS0) {
#pragma unroll
                            for (int mm = 0; mm < MB; ++mm)
                                sb[(mm * L0 + u) * NCOL + col] = Lin<MODE, TOUT>::put(cswap(x0[it][j * MB + mm]));
                        }
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if ((tid & 31) == 0) mbar_arrive(bar_stg_full + 8 * kbuf);
                    if constexpr (!SEP) {
                        if (tid == 0) producer_store_round(tile, j);
                    }
                    ++rnd;
                }
            } else {
#pragma unroll
                for (int it = 0; it < ITERS0; ++it) {
                    const int w = tid + it * NT;
                    int col, u;
                    map0(it, col, u);
                    if (ITEMS0 % NT == 0 || w < ITEMS0) {
                        TOUT *op = ob + (long long)u * p.o_sn + col * CH;
                        if (full_tile) {
#pragma unroll
                            for (int m = 0; m < R0; ++m) GIO<MODE, TOUT>::store(op + (long long)(m * L0) * p.o_sn, cswap(x0[it][m]));
                        } else {
                            const bool colok = (ce0 + col) < CE;
#pragma unroll
                            for (int m = 0; m < R0; ++m)
                                if (colok && u + m * L0 < p.n_out) GIO<MODE, TOUT>::store(op + (long long)(m * L0) * p.o_sn, cswap(x0[it][m]));
                        }
                    }
                }
            }
        }
        if constexpr (!TMA_IN) cta_sync<NT, SEP>();  // the next tile's stage 0 overwrites the buffer and the gate tables
        SPX_MARK(7)
        ++tl_tile;
        ++tile_it;
    }
#undef SPX_MARK
    if constexpr (TMA_IN && !SEP) {
        if (tid == 0) tma_wait_all();   // staging buffers must outlive the last TMA stores
    }
    }   // compute threads
    if constexpr (TMEM_IO) {
        tc_fence_before();
        __syncthreads();                // every TMEM access of this CTA is done
        tc_fence_after();
        if (tid >= NT && tid < NT + 32) tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace spx
