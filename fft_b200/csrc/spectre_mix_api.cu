// spectre_mix_api.cu -- C ABI (include/spectre_mix.h), plan selection, twiddle tables, launches.
//
// Host-side counterpart of the reference's call sites spectre.py:506, :542-553
// (forward mix) and :776-777 (prefill rfft).  No torch types, no cuFFT, no CPU
// fallback: anything this file cannot run returns an error code.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/spectre_mix.h"
#include "spectre_internal.h"
#include "spectre_long.h"
#include "spectre_mix_registry.h"

namespace spx {
namespace {
thread_local std::string g_err;
}
// error reporting shared by every translation unit of the library (spectre_internal.h)
int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
int cuda_fail(cudaError_t e, const char *what) {
    cudaGetLastError();  // clear the sticky-less error state
    return fail(SPECTRE_MIX_ERR_CUDA + (int)e, "%s: %s", what, cudaGetErrorString(e));
}
const char *last_error() { return g_err.c_str(); }
}  // namespace spx

namespace {

using spx::cuda_fail;
using spx::fail;
using spx::KernelEntry;
using spx::MixParams;

// Tuning knobs (experiments / diagnostics only): process-wide atomics, read once per call with relaxed loads, so a call
// always sees one consistent value of each and concurrent setters are not a data race.
std::atomic<int> g_tile_channels_override{0};
std::atomic<int> g_prefetch{0};   // L2 prefetch of the next tiles: 0 off (default: the TMA unit is the busiest part of the TMEM variant), 1 TMA prefetch, 2 cooperative whole-line prefetch (needs -DSPX_COOP_PF=1)
std::atomic<int> g_use_tma{1};
std::atomic<int> g_use_tmem{1};
std::atomic<int> g_skew_ns{-350};   // warp stagger after the CTA barriers (see stagger() in the kernel header)
std::atomic<int> g_sched{3};        // bit 0 stagger before the last inverse pass too, bit 1 split barrier around its read
std::atomic<int> g_use_two_pass{1};
std::atomic<int> g_long_chunk_mb{0};  // long-context path: rows per chunk sized so the intermediate stays in L2 (0 = whole batch at once)
std::atomic<int> g_l2_promo{0};       // L2 promotion of the input tensor map: 0 none, 1 64 B, 2 128 B, 3 256 B
std::atomic<unsigned long long *> g_timeline{nullptr};

// ---------------------------------------------------------------- kernel registry
const std::vector<KernelEntry> &registry() {
    static std::vector<KernelEntry> all = [] {
        std::vector<KernelEntry> v;
        typedef const KernelEntry *(*TableFn)(int *);
        TableFn fns[] = {spx::table_small, spx::table_1024, spx::table_2048,
                         spx::table_4096,  spx::table_8192, spx::table_16384};
        for (TableFn f : fns) {
            int n = 0;
            const KernelEntry *t = f(&n);
            v.insert(v.end(), t, t + n);
        }
        return v;
    }();
    return all;
}

int mode_channels(int mode) { return mode == spx::MODE_QUAD ? 4 : (mode == spx::MODE_PAIR ? 2 : 1); }

// ---------------------------------------------------------------- per-device state
// Per-device state.  Nothing here is touched on the launch path except through lock-free reads: the device attributes are
// filled once (std::call_once), a twiddle table is published through an atomic pointer once built (build + upload under the
// device's mutex, first use of an n_fft only), the occupancy cache is a short critical section of its own.  Kernel launches
// themselves run outside every lock, so host threads driving different streams never serialise on the library.
constexpr int kMaxDevices = 64;
constexpr int kMaxLog2N = 15;
struct DeviceState {
    std::once_flag once;
    int init_rc = 0;
    cudaError_t init_err = cudaSuccess;
    int sm_count = 0;
    int max_smem_optin = 0;
    std::atomic<float2 *> twiddles[3 * (kMaxLog2N + 1)] = {};   // (log2(n_fft), radix order) -> device table (plain and SUB plans share it)
    std::atomic<float2 *> long_tw[5] = {};                // R (2, 4) -> twiddles of the long-context streaming passes
    std::mutex mu;                                        // table creation, occupancy cache, pool creation
    std::map<std::pair<const KernelEntry *, int>, int> occupancy;  // (entry, flags) -> CTAs/SM
    // Scratch of the two-pass long-context path when the caller passes no workspace: stream-ordered allocations
    // (cudaMallocFromPoolAsync / cudaFreeAsync on the CALLER's stream) from a private pool that keeps its memory, so calls on
    // different streams can never share a buffer, nothing synchronises the device, and the pair is legal under stream capture.
    std::atomic<cudaMemPool_t> pool{nullptr};
};
DeviceState g_dev[kMaxDevices];

// log2(x) when x is a power of two, else -1
int ilog2_exact(int x) {
    if (x <= 0 || (x & (x - 1))) return -1;
    int s = 0;
    while ((1 << s) < x) ++s;
    return s;
}

// W_{N/P(s)}^{u q} for every non-final stage, laid out exactly as Plan::TWOFF expects.
std::vector<float2> build_twiddles(const int radix[4]) {
    int N = radix[0] * radix[1] * radix[2] * radix[3];
    int ns = radix[3] > 1 ? 4 : (radix[2] > 1 ? 3 : 2);
    std::vector<float2> t;
    int P = 1;
    for (int s = 0; s < ns - 1; ++s) {
        const int R = radix[s], L = N / (P * R);
        // radix-16 stages keep rows q = 1, 2, 3, 4, 8, 12 only (spx::tw_rows): the kernel multiplies the others together
        // (SPX_TW4 builds: q = 1, 2, 4, 8 for the long tables, spx::tw_rows4)
        static const int rows16[6] = {1, 2, 3, 4, 8, 12};
        static const int rows16_4[4] = {1, 2, 4, 8};
        const int nrows = spx::tw_rows(R, L);
        for (int j = 0; j < nrows; ++j)
            for (int u = 0; u < L; ++u) {
                const int q = (R == 16) ? (spx::tw_rows4(R, L) ? rows16_4[j] : rows16[j]) : j + 1;
                // exponent reduced modulo the period before scaling keeps the angle exact in double
                const long long period = N / P;
                const long long e = ((long long)u * q) % period;
                const double ang = -2.0 * M_PI * (double)e / (double)period;
                t.push_back(make_float2((float)cos(ang), (float)sin(ang)));
            }
        P *= R;
    }
    return t;
}

int get_device_state(DeviceState **out, int *dev_out) {
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(SPECTRE_MIX_ERR_NO_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
    }
    if (dev < 0 || dev >= kMaxDevices) return fail(SPECTRE_MIX_ERR_NO_DEVICE, "device ordinal %d out of range", dev);
    DeviceState &st = g_dev[dev];
    std::call_once(st.once, [&] {
        st.init_err = cudaDeviceGetAttribute(&st.sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (st.init_err == cudaSuccess)
            st.init_err = cudaDeviceGetAttribute(&st.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    });
    if (st.init_err != cudaSuccess) return cuda_fail(st.init_err, "cudaDeviceGetAttribute");
    *out = &st;
    if (dev_out) *dev_out = dev;
    return 0;
}

int get_twiddles(DeviceState &st, const KernelEntry &k, const float2 **tw) {
    int lg = 0;
    while ((1 << lg) < k.n_fft) ++lg;
    // radix orders of one n_fft: (r, 16, 16 ..) | (16, 2, 16, 16) at 8192 | (16, 16, r) at 1024 / 2048: one table each
    const int last = k.radix[3] > 1 ? k.radix[3] : (k.radix[2] > 1 ? k.radix[2] : k.radix[1]);
    if (k.radix[0] == 16 && k.radix[1] == 2) lg += kMaxLog2N + 1;
    else if (k.radix[0] == 16 && last != 16) lg += 2 * (kMaxLog2N + 1);
    float2 *d = st.twiddles[lg].load(std::memory_order_acquire);
    if (!d) {
        // first use of this n_fft on this device: build in double on the host, upload (synchronous -- do one warm-up call
        // before capturing a CUDA graph), publish.  Later calls take the lock-free path above.
        std::lock_guard<std::mutex> lock(st.mu);
        d = st.twiddles[lg].load(std::memory_order_acquire);
        if (!d) {
            std::vector<float2> h = build_twiddles(k.radix);
            if ((int)h.size() != k.twn) return fail(SPECTRE_MIX_ERR_BAD_ARG, "internal: twiddle count %zu != %d", h.size(), k.twn);
            cudaError_t e = cudaMalloc(&d, h.size() * sizeof(float2));
            if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(twiddles)");
            e = cudaMemcpy(d, h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) { cudaFree(d); return cuda_fail(e, "cudaMemcpy(twiddles)"); }
            st.twiddles[lg].store(d, std::memory_order_release);
        }
    }
    *tw = d;
    return 0;
}

// W_{R 4096}^{u q}, q = 1 .. R-1, u < 4096, for the pre / post passes of the three-launch long-context path
int get_long_twiddles(DeviceState &st, int R, const float2 **tw) {
    float2 *d = st.long_tw[R].load(std::memory_order_acquire);
    if (!d) {
        std::lock_guard<std::mutex> lock(st.mu);
        d = st.long_tw[R].load(std::memory_order_acquire);
        if (!d) {
            const int sub = 4096;
            const long long N = (long long)R * sub;
            std::vector<float2> h((size_t)(R - 1) * sub);
            for (int q = 1; q < R; ++q)
                for (int u = 0; u < sub; ++u) {
                    const double ang = -2.0 * M_PI * (double)(((long long)u * q) % N) / (double)N;
                    h[(size_t)(q - 1) * sub + u] = make_float2((float)cos(ang), (float)sin(ang));
                }
            cudaError_t e = cudaMalloc(&d, h.size() * sizeof(float2));
            if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc(long-context twiddles)");
            e = cudaMemcpy(d, h.data(), h.size() * sizeof(float2), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) { cudaFree(d); return cuda_fail(e, "cudaMemcpy(long-context twiddles)"); }
            st.long_tw[R].store(d, std::memory_order_release);
        }
    }
    *tw = d;
    return 0;
}

// private stream-ordered pool of a device (created on first use, keeps what it allocated: release threshold = max)
int get_pool(DeviceState &st, int dev, cudaMemPool_t *out) {
    cudaMemPool_t p = st.pool.load(std::memory_order_acquire);
    if (!p) {
        std::lock_guard<std::mutex> lock(st.mu);
        p = st.pool.load(std::memory_order_acquire);
        if (!p) {
            cudaMemPoolProps props;
            memset(&props, 0, sizeof(props));
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            cudaError_t e = cudaMemPoolCreate(&p, &props);
            if (e != cudaSuccess) return cuda_fail(e, "cudaMemPoolCreate(long-context scratch pool)");
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(p, cudaMemPoolAttrReleaseThreshold, &keep);
            st.pool.store(p, std::memory_order_release);
        }
    }
    *out = p;
    return 0;
}

// ---------------------------------------------------------------- plan selection
struct Choice {
    const KernelEntry *k = nullptr;
    int gate_tables = 1;
    int tiles_per_row = 0;
    size_t smem = 0;
};

bool aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// Widest element mode the layout allows: QUAD needs 4 | group_width and 16-byte (fp32) / 8-byte (bf16)
// aligned rows, PAIR needs 2 | group_width, REAL always works.
int pick_mode(int dtype, int group_width, const void *v, long long v_sb, long long v_sn, const void *out,
              long long o_sb, long long o_sn, const void *mem, long long mem_stride) {
    const size_t es = dtype == SPECTRE_MIX_F32 ? 4 : 2;
    auto ok = [&](int ch) {
        if (group_width % ch) return false;
        const size_t a = es * ch;
        if (v && !aligned(v, a)) return false;
        if (out && !aligned(out, a)) return false;
        if ((v_sb % ch) || (v_sn % ch) || (o_sb % ch) || (o_sn % ch)) return false;
        if (mem && ch > 1 && (!aligned(mem, 16) || (mem_stride % 2))) return false;
        return true;
    };
    if (ok(4)) return spx::MODE_QUAD;
    if (ok(2)) return spx::MODE_PAIR;
    return spx::MODE_REAL;
}

// B > 0: batch-aware choice between the wide-row TMEM-staged variants of n_fft = 1024 / 2048 (128 KB tiles, one CTA per SM)
// and the narrower two-CTAs-per-SM variants: below about seven tiles per SM the launch is too short for the big tiles
// (measured crossover: profiles/r02k_ab_wide_rows_*.txt -- 1024: batch 48, 2048: batch 16-24 at C = 768)
int choose(const DeviceState &st, int n_fft, int dtype, int mode_max, int C, int group_width, Choice *out,
           bool no_gate = false, int B = 0) {
    const std::vector<KernelEntry> &reg = registry();
    const int g_tile_channels_override = ::g_tile_channels_override.load(std::memory_order_relaxed);
    // candidates in registry order (first = default) for the widest mode that has any variant
    for (int mode = mode_max; mode <= spx::MODE_REAL; ++mode) {
        const KernelEntry *best = nullptr;
        Choice bc;
        for (const KernelEntry &k : reg) {
            if (k.n_fft != n_fft || k.io != dtype || k.mode != mode || k.sub || k.dit) continue;
            const int ch = mode_channels(mode);
            const int tw = ch * k.ncol;  // channels per tile
            if (g_tile_channels_override && tw != g_tile_channels_override) continue;
            // gate groups a tile can touch (tiles start at multiples of tw)
            int gt;
            if (group_width % tw == 0) gt = 1;
            else if (tw % group_width == 0) gt = tw / group_width;
            else gt = (tw + group_width - 1) / group_width + 1;
            if (no_gate) gt = 0;
            const size_t sm = k.smem_bytes(gt, false, false);
            if ((int)sm > st.max_smem_optin) continue;
            const int ce = C / ch;
            Choice c;
            c.k = &k;
            c.gate_tables = gt;
            c.tiles_per_row = (ce + k.ncol - 1) / k.ncol;
            c.smem = sm;
            if (k.tmem_ok && n_fft < 4096 && B > 0 && !g_tile_channels_override &&
                (long long)B * c.tiles_per_row < 7LL * st.sm_count)
                continue;   // short launch: a narrower variant of this n_fft follows in the registry
            if (!best) { best = &k; bc = c; }
        }
        if (best) { *out = bc; return 0; }
        if (g_tile_channels_override) continue;
    }
    return fail(SPECTRE_MIX_ERR_UNSUPPORTED,
                "no kernel variant for n_fft=%d dtype=%d (supported: powers of two in [32, 16384]%s)", n_fft, dtype,
                g_tile_channels_override ? "; tile override active" : "");
}

long long algorithmic_bytes(int dtype, bool has_mem, int B, int N, int n_fft, int C, int group_width) {
    const long long es = dtype == SPECTRE_MIX_F32 ? 4 : 2;
    const long long n_io = std::min(N, n_fft), fh = n_fft / 2 + 1, ng = C / group_width;
    return (long long)B * n_io * C * es * 2 + (long long)B * ng * fh * 8 + (has_mem ? fh * C * 8 : 0);
}

int check_common(int B, int N, int n_fft, int C, int group_width) {
    if (B < 0 || N < 0 || C < 0) return fail(SPECTRE_MIX_ERR_BAD_ARG, "negative size (B=%d N=%d C=%d)", B, N, C);
    if (group_width <= 0 || (C % group_width) != 0)
        return fail(SPECTRE_MIX_ERR_BAD_ARG, "C=%d is not a multiple of group_width=%d", C, group_width);
    if (n_fft < 32 || n_fft > 16384 || (n_fft & (n_fft - 1)))
        return fail(SPECTRE_MIX_ERR_UNSUPPORTED, "n_fft=%d unsupported: need a power of two in [32, 16384]", n_fft);
    return 0;
}

int occupancy_of(DeviceState &st, const Choice &c, bool has_mem, bool tma, bool tmem = false) {
    auto key = std::make_pair(c.k, c.gate_tables * 8 + (has_mem ? 1 : 0) + (tma ? 2 : 0) + (tmem ? 4 : 0));
    std::lock_guard<std::mutex> lock(st.mu);   // cache lookup only; the launch itself runs outside the lock
    auto it = st.occupancy.find(key);
    if (it != st.occupancy.end()) return it->second;
    int occ = c.k->occupancy(c.gate_tables, has_mem, tma, tmem);
    st.occupancy[key] = occ;
    return occ;
}

// ---------------------------------------------------------------- TMA descriptor for V ([B][rows][C], channels contiguous)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            p = nullptr;
        }
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

// true when V can be described to the TMA unit: 16-byte aligned base and strides, extents within 32 bits
bool tma_layout_ok(const void *v, int dtype, long long v_sb, long long v_sn) {
    const long long es = dtype == SPECTRE_MIX_F32 ? 4 : 2;
    return aligned(v, 16) && (v_sb * es) % 16 == 0 && (v_sn * es) % 16 == 0 && v_sn > 0 && v_sb > 0;
}

bool make_v_tensor_map(CUtensorMap *tm, const void *v, int dtype, long long v_sb, long long v_sn, int B, int rows, int C,
                       int box_rows, int tile_channels, int l2_promo = 0) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return false;
    const cuuint64_t es = dtype == SPECTRE_MIX_F32 ? 4 : 2;
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)rows, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)v_sn * es, (cuuint64_t)v_sb * es};
    cuuint32_t box[3] = {(cuuint32_t)tile_channels, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(tm, dtype == SPECTRE_MIX_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3,
                     const_cast<void *>(v), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_NONE,
                     l2_promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                   : (l2_promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                    : (l2_promo == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE)),
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}


// DIT2 variant: V / out seen as [B][n][parity][C] (row = 2 n + parity), box {tile channels, 2, box_rows, 1}
bool make_dit_tensor_map(CUtensorMap *tm, const void *v, int dtype, long long v_sb, long long v_sn, int B, int rows, int C,
                         int box_rows, int tile_channels, int l2_promo = 0) {
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc || (rows & 1)) return false;
    const cuuint64_t es = dtype == SPECTRE_MIX_F32 ? 4 : 2;
    cuuint64_t dims[4] = {(cuuint64_t)C, 2, (cuuint64_t)(rows / 2), (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)v_sn * es, (cuuint64_t)v_sn * es * 2, (cuuint64_t)v_sb * es};
    cuuint32_t box[4] = {(cuuint32_t)tile_channels, 2, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = enc(tm, dtype == SPECTRE_MIX_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4,
                     const_cast<void *>(v), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                     l2_promo == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                   : (l2_promo == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                                    : (l2_promo == 3 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE)),
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// the DIT2 kernel of a transform length (n_fft = 2 * its sub-transform), if this call can take it: packed layout, even row
// count (rows pair up as (2 n, 2 n + 1)), TMA + TMEM staging on, not switched off (sched bit 5 = experiments)
const KernelEntry *dit_kernel(const DeviceState &st, int n_fft, int dtype, int mode, int n_io, const void *mem, long long mem_stride) {
    if (mode != spx::MODE_QUAD || (n_io & 1) || g_use_two_pass.load(std::memory_order_relaxed) == 2) return nullptr;
    if (!g_use_tma.load(std::memory_order_relaxed) || !g_use_tmem.load(std::memory_order_relaxed)) return nullptr;
    if (g_tile_channels_override.load(std::memory_order_relaxed) || (g_sched.load(std::memory_order_relaxed) & 32)) return nullptr;
    if (mem && (mem_stride % 2 != 0)) return nullptr;
    for (const KernelEntry &k : registry())
        if (k.dit == 2 && 2 * k.n_fft == n_fft && k.io == dtype && k.mode == spx::MODE_QUAD && k.tmem_ok &&
            (int)k.smem_bytes(2, true, true) <= st.max_smem_optin)
            return &k;
    return nullptr;
}

int mix_dit(DeviceState &st, const KernelEntry &k, const void *v, int dtype, long long v_sb, long long v_sn, const void *gate,
            const spx::GateSrc *gs, const void *mem, long long mem_stride, void *out, long long o_sb, long long o_sn, int B, int n_io,
            int n_fft, int C, int group_width, cudaStream_t stream) {
    const float2 *tw = nullptr;
    if (int rc = get_twiddles(st, k, &tw)) return rc;
    MixParams p;
    memset(&p, 0, sizeof(p));
    p.v = v;
    p.out = out;
    p.gate = reinterpret_cast<const float2 *>(gate);
    if (gs) p.gsrc = *gs;
    p.mem = reinterpret_cast<const float2 *>(mem);
    p.tw = tw;
    p.v_sb = v_sb; p.v_sn = v_sn; p.o_sb = o_sb; p.o_sn = o_sn;
    p.mem_stride = mem_stride;
    p.B = B;
    p.n_in = p.n_out = n_io / 2;               // rows of each of the two interleaved sub-sequences
    p.C = C;
    p.group_width = group_width;
    p.NG = C / group_width;
    p.tiles_per_row = C / 4;                   // one 4-channel group per tile, its even and odd rows are the two tile columns
    p.num_tiles = B * p.tiles_per_row;
    p.gate_tables = 2;                         // half spectrum of the 2 N-point transform = two table slots
    p.inv_n = 1.0f / (float)n_fft;
    p.prefetch = 0;
    p.skew_ns = g_skew_ns.load(std::memory_order_relaxed);
    p.sched = g_sched.load(std::memory_order_relaxed) & ~(16 | 8 | 4);
    p.timeline = g_timeline.load(std::memory_order_relaxed);
    p.sub_R = 1;
    p.gw_shift = ilog2_exact(group_width);
    alignas(64) CUtensorMap tmap, tmap_out;
    if (!make_dit_tensor_map(&tmap, v, dtype, v_sb, v_sn, B, n_io, C, k.tmem_box_rows, 4, g_l2_promo.load(std::memory_order_relaxed)) ||
        !make_dit_tensor_map(&tmap_out, out, dtype, o_sb, o_sn, B, n_io, C, k.tmem_box_rows, 4))
        return fail(SPECTRE_MIX_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled (rank 4) failed");
    const int grid = std::min(p.num_tiles, st.sm_count);
    cudaError_t e = k.launch(p, grid, mem != nullptr, &tmap, &tmap_out, true, stream);
    if (e != cudaSuccess) return cuda_fail(e, "DIT2 kernel launch");
    return 0;
}

// ---------------------------------------------------------------- long-context two-pass path (n_fft = 8192, 16384)
const KernelEntry *find_sub_kernel() {
    for (const KernelEntry &k : registry())
        if (k.sub && k.n_fft == 4096 && k.mode == spx::MODE_QUAD && k.io == SPECTRE_MIX_F32) return &k;
    return nullptr;
}

// pre pass (radix-R stage + twiddles) -> 4096-point shared-memory kernel on R interleaved sub-transforms (in place in
// the scratch tensor) -> post pass.  Three launches, each streaming the tensor once.
int mix_two_pass_rows(DeviceState &st, const KernelEntry &k, const void *v, int dtype, long long v_sb, long long v_sn, const void *gate,
                      const spx::GateSrc *gs, const void *mem, long long mem_stride, void *out, long long o_sb, long long o_sn, int B,
                      int n_io, int n_fft, int C, int group_width, float *scr, cudaStream_t stream) {
    const int sub = 4096, R = n_fft / sub;
    const float2 *ltw = nullptr;
    if (int rc = get_long_twiddles(st, R, &ltw)) return rc;
    cudaError_t e = spx::long_pass(true, R, dtype == SPECTRE_MIX_BF16, v, scr, v_sb, v_sn, B, n_io, C, sub, st.sm_count, ltw, stream);
    if (e != cudaSuccess) return cuda_fail(e, "long-context pre pass");

    const float2 *tw = nullptr;
    if (int rc = get_twiddles(st, k, &tw)) return rc;
    MixParams p;
    memset(&p, 0, sizeof(p));
    p.v = scr;
    p.out = scr;
    p.gate = reinterpret_cast<const float2 *>(gate);
    if (gs) p.gsrc = *gs;                    // gate == nullptr: the sub-transform kernel evaluates the gate from the anchors
    p.mem = reinterpret_cast<const float2 *>(mem);
    p.tw = tw;
    p.v_sb = p.o_sb = (long long)sub * C;
    p.v_sn = p.o_sn = C;
    p.mem_stride = mem_stride;
    p.B = B * R;
    p.n_in = p.n_out = sub;
    p.C = C;
    p.group_width = group_width;
    p.NG = C / group_width;
    const int tile_ch = mode_channels(k.mode) * k.ncol;
    p.tiles_per_row = (C / 4 + k.ncol - 1) / k.ncol;
    p.num_tiles = B * R * p.tiles_per_row;
    p.gate_tables = 2;                       // one full-length table = two half-length slots
    p.inv_n = 1.0f / (float)n_fft;
    p.prefetch = g_prefetch.load(std::memory_order_relaxed);
    p.timeline = nullptr;
    p.skew_ns = g_skew_ns.load(std::memory_order_relaxed);
    p.sched = g_sched.load(std::memory_order_relaxed);
    p.sub_R = R;
    p.sub_shift = ilog2_exact(R);
    p.gw_shift = ilog2_exact(group_width);
    Choice c;
    c.k = &k;
    c.gate_tables = 2;
    alignas(64) CUtensorMap tmap, tmap_out;
    const bool use_tma = g_use_tma.load(std::memory_order_relaxed) != 0, use_tmem = g_use_tmem.load(std::memory_order_relaxed) != 0;
    bool tma = use_tma && k.tma_ok && (int)k.smem_bytes(2, true, false) <= st.max_smem_optin &&
               make_v_tensor_map(&tmap, scr, SPECTRE_MIX_F32, p.v_sb, p.v_sn, B * R, sub, C, spx::kTmaBoxRows, tile_ch,
                                 g_l2_promo.load(std::memory_order_relaxed)) &&
               make_v_tensor_map(&tmap_out, scr, SPECTRE_MIX_F32, p.o_sb, p.o_sn, B * R, sub, C, k.out_box_rows, tile_ch);
    const bool tmem = tma && use_tmem && k.tmem_ok && (int)k.smem_bytes(2, true, true) <= st.max_smem_optin;
    if (!tma && (int)k.smem_bytes(2, false, false) > st.max_smem_optin)
        return fail(SPECTRE_MIX_ERR_UNSUPPORTED, "internal: sub-transform kernel does not fit shared memory");
    const int occ = std::max(1, occupancy_of(st, c, mem != nullptr, tma, tmem));
    p.pair_tiles = (tmem && (p.sched & 16) && p.tiles_per_row % 2 == 0 && group_width % (2 * tile_ch) == 0) ? 1 : 0;
    const int grid = std::min(p.pair_tiles ? p.num_tiles / 2 : p.num_tiles, st.sm_count * occ);
    e = k.launch(p, grid, mem != nullptr, tma ? &tmap : nullptr, tma ? &tmap_out : nullptr, tmem, stream);
    if (e != cudaSuccess) return cuda_fail(e, "long-context sub-transform kernel");

    e = spx::long_pass(false, R, dtype == SPECTRE_MIX_BF16, scr, out, o_sb, o_sn, B, n_io, C, sub, st.sm_count, ltw, stream);
    if (e != cudaSuccess) return cuda_fail(e, "long-context post pass");
    return 0;
}

// rows of the batch one round of the three passes takes: the whole batch unless the L2-chunk experiment knob is set
int two_pass_rows(int B, int n_fft, int C) {
    const size_t row_bytes = (size_t)n_fft * C * sizeof(float);
    size_t chunk_mb = (size_t)g_long_chunk_mb.load(std::memory_order_relaxed);
    if (const char *env = getenv("SPECTRE_MIX_LONG_CHUNK_MB")) chunk_mb = (size_t)std::max(0L, strtol(env, nullptr, 10));   // experiment knob
    return chunk_mb == 0 ? B : (int)std::max<size_t>(1, std::min<size_t>((size_t)B, (chunk_mb << 20) / std::max<size_t>(row_bytes, 1)));
}

// does this call take the two-pass long-context path?  (QUAD layout, group width a multiple of 8, n_fft > 4096)
// g_use_two_pass: 0 never, 1 automatic (default: only where no TMEM-staged single-kernel variant of this n_fft exists -- the
// single pass over HBM wins: 2600 against 1640 GB/s at n_fft = 8192, profiles/r02f_8192_single_kernel.txt), 2 wherever possible
const KernelEntry *two_pass_kernel(int n_fft, int dtype, int mode, int group_width, const void *mem, long long mem_stride) {
    const int knob = g_use_two_pass.load(std::memory_order_relaxed);
    if (!knob || n_fft <= 4096 || mode != spx::MODE_QUAD || group_width % 8 != 0) return nullptr;
    if (mem && (mem_stride % 2 != 0)) return nullptr;
    if (knob == 1 && g_use_tma.load(std::memory_order_relaxed) && g_use_tmem.load(std::memory_order_relaxed) &&
        !g_tile_channels_override.load(std::memory_order_relaxed)) {
        // the packed layout is 16-byte aligned by construction, so a TMEM-staged variant can always take it
        for (const KernelEntry &k : registry())
            if (k.n_fft == n_fft && k.io == dtype && k.mode == spx::MODE_QUAD && !k.sub && !k.dit) {
                if (k.tmem_ok && group_width % (4 * k.ncol) == 0) return nullptr;   // the default (first) packed variant is TMEM-staged
                break;
            }
    }
    return find_sub_kernel();
}

// The complex intermediate [rows][R][4096][C] fp32 lives in `scr` (the caller's workspace or a stream-ordered allocation made
// for this call, see mix_fwd_impl).  With the chunk knob the batch is walked in chunks of rows whose intermediate is small
// enough to stay in the 126 MB L2 between the three launches.
int mix_two_pass(DeviceState &st, const KernelEntry &k, const void *v, int dtype, long long v_sb, long long v_sn,
                 const void *gate, const spx::GateSrc *gs, const void *mem, long long mem_stride, void *out, long long o_sb,
                 long long o_sn, int B, int n_io, int n_fft, int C, int group_width, float *scr, cudaStream_t stream) {
    const int rows = two_pass_rows(B, n_fft, C);
    const size_t es = dtype == SPECTRE_MIX_F32 ? 4 : 2;
    const int NG = C / group_width;
    const size_t gate_row = (size_t)NG * (n_fft / 2 + 1) * sizeof(float2);
    int rc = 0;
    for (int b0 = 0; b0 < B && !rc; b0 += rows) {
        const int nb = std::min(rows, B - b0);
        spx::GateSrc gsub;
        if (gs) {   // anchors / phase of this chunk of batch rows
            gsub = *gs;
            gsub.anchors += (size_t)b0 * NG * gs->Bk;
            if (gsub.pos) gsub.pos += (size_t)b0 * gs->pos_stride_b;
        }
        rc = mix_two_pass_rows(st, k, reinterpret_cast<const char *>(v) + (size_t)b0 * v_sb * es, dtype, v_sb, v_sn,
                               gate ? reinterpret_cast<const char *>(gate) + (size_t)b0 * gate_row : nullptr, gs ? &gsub : nullptr,
                               mem, mem_stride, reinterpret_cast<char *>(out) + (size_t)b0 * o_sb * es, o_sb, o_sn, nb, n_io, n_fft,
                               C, group_width, scr, stream);
    }
    return rc;
}

}  // namespace

extern "C" {

int spectre_mix_abi_version(void) { return 2; }

const char *spectre_mix_last_error(void) { return spx::last_error(); }

int spectre_mix_set_tile_channels(int tile_channels) {
    if (tile_channels < 0) return fail(SPECTRE_MIX_ERR_BAD_ARG, "tile_channels < 0");
    g_tile_channels_override.store(tile_channels);
    return 0;
}

int spectre_mix_set_prefetch(int enable) {
    g_prefetch.store(enable < 0 ? 0 : enable);   // 0 off, 1 TMA prefetch of the next tiles, 2 cooperative whole-line prefetch (TMEM variant)
    return 0;
}

int spectre_mix_set_timeline(void *device_buffer) {
    g_timeline.store(reinterpret_cast<unsigned long long *>(device_buffer));
    return 0;
}

int spectre_mix_set_two_pass(int enable) {
    g_use_two_pass.store(enable < 0 ? 0 : (enable > 2 ? 2 : enable));
    return 0;
}

int spectre_mix_set_skew_ns(int code) {
    g_skew_ns.store(code);
    return 0;
}

int spectre_mix_set_l2_promotion(int level) {
    g_l2_promo.store(level < 0 ? 0 : (level > 3 ? 3 : level));
    return 0;
}

int spectre_mix_set_sched(int flags) {
    g_sched.store(flags);
    return 0;
}

int spectre_mix_set_tmem(int enable) {
    g_use_tmem.store(enable ? 1 : 0);
    return 0;
}

int spectre_mix_set_tma(int enable) {
    g_use_tma.store(enable ? 1 : 0);
    return 0;
}

namespace {
size_t align256(size_t x) { return (x + 255) / 256 * 256; }

// `gs` != nullptr: the gate comes from anchors (spectre_mix_fwd_anchors) and `gate` is null
int mix_fwd_impl(const void *v, int v_dtype, int64_t v_stride_b, int64_t v_stride_n, const void *gate, const spx::GateSrc *gs,
                 const void *mem, int64_t mem_stride, void *out, int out_dtype, int64_t out_stride_b, int64_t out_stride_n, int B,
                 int N, int n_fft, int C, int group_width, void *ws, size_t ws_bytes, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (int rc = check_common(B, N, n_fft, C, group_width)) return rc;
    if (v_dtype != out_dtype || (v_dtype != SPECTRE_MIX_F32 && v_dtype != SPECTRE_MIX_BF16))
        return fail(SPECTRE_MIX_ERR_UNSUPPORTED, "dtype pair (%d, %d) unsupported: V and out must both be f32 or both bf16",
                    v_dtype, out_dtype);
    const int n_io = std::min(N, n_fft);
    if (B == 0 || C == 0 || n_io == 0) return 0;  // empty input: nothing to write
    if (!v || (!gate && !gs) || !out) return fail(SPECTRE_MIX_ERR_BAD_ARG, "null pointer (v=%p gate=%p out=%p)", v, gate, out);
    if (gate && !aligned(gate, 8)) return fail(SPECTRE_MIX_ERR_UNSUPPORTED, "gate must be 8-byte aligned");
    if (mem && !aligned(mem, 8)) return fail(SPECTRE_MIX_ERR_UNSUPPORTED, "memory must be 8-byte aligned");
    if (mem && mem_stride < C) return fail(SPECTRE_MIX_ERR_BAD_ARG, "mem_stride=%lld < C=%d", (long long)mem_stride, C);
    const int NG = C / group_width, F_half = n_fft / 2 + 1;
    if (gs) {
        if (!gs->anchors || !gs->bias || !gs->eps) return fail(SPECTRE_MIX_ERR_BAD_ARG, "null pointer (anchors / bias / eps)");
        if (gs->G <= 0 || NG % gs->G != 0 || gs->Bk < 1)
            return fail(SPECTRE_MIX_ERR_BAD_ARG, "anchors: NG=%d is not a multiple of G=%d, or Bk=%d < 1", NG, gs->G, gs->Bk);
        if (!aligned(gs->anchors, 8) || (gs->pos && !aligned(gs->pos, 8))) return fail(SPECTRE_MIX_ERR_UNSUPPORTED, "anchors / phase must be 8-byte aligned");
    }

    DeviceState *st = nullptr;
    int dev = 0;
    if (int rc = get_device_state(&st, &dev)) return rc;

    const int mode = pick_mode(v_dtype, group_width, v, v_stride_b, v_stride_n, out, out_stride_b, out_stride_n, mem,
                               mem_stride);
    const bool layout_tma = tma_layout_ok(v, v_dtype, v_stride_b, v_stride_n) && tma_layout_ok(out, out_dtype, out_stride_b, out_stride_n);
    const KernelEntry *kd = (n_fft > 4096 && layout_tma) ? dit_kernel(*st, n_fft, v_dtype, mode, n_io, mem, mem_stride) : nullptr;
    const KernelEntry *ks = kd ? nullptr : two_pass_kernel(n_fft, v_dtype, mode, group_width, mem, mem_stride);
    Choice c;
    if (kd) {
        c.k = kd;
    } else if (!ks) {
        if (int rc = choose(*st, n_fft, v_dtype, mode, C, group_width, &c, false, B)) return rc;
    }
    // scratch of this call: [long-context intermediate][materialised gate, when this layout has no in-kernel gate generator]
    const bool fused_gate = gs && (ks ? ks->anch_ok : c.k->anch_ok);
    const size_t need_long = ks ? (size_t)two_pass_rows(B, n_fft, C) * (size_t)n_fft * C * sizeof(float) : 0;
    const size_t need_gate = (gs && !fused_gate) ? (size_t)B * NG * F_half * sizeof(float2) : 0;
    const size_t need = align256(need_long) + need_gate;
    char *scr = nullptr;
    bool pooled = false;
    if (need) {
        if (ws) {
            if (ws_bytes < need)
                return fail(SPECTRE_MIX_ERR_BAD_ARG, "workspace too small: %zu bytes given, %zu needed (spectre_mix%s_workspace_bytes)",
                            ws_bytes, need, gs ? "_anchors" : "");
            if (!aligned(ws, 16)) return fail(SPECTRE_MIX_ERR_UNSUPPORTED, "workspace must be 16-byte aligned");
            scr = reinterpret_cast<char *>(ws);
        } else {
            // no workspace passed: a stream-ordered allocation from the device's private pool on THIS call's stream -- two calls
            // on different streams never share a buffer (the round-1 per-device scratch did, and raced), nothing synchronises
            // the device, and the pair alloc / free is legal under stream capture
            cudaMemPool_t pool = nullptr;
            if (int rc = get_pool(*st, dev, &pool)) return rc;
            void *pm = nullptr;
            cudaError_t e = cudaMallocFromPoolAsync(&pm, need, pool, stream);
            if (e != cudaSuccess) return cuda_fail(e, "cudaMallocFromPoolAsync(scratch)");
            scr = reinterpret_cast<char *>(pm);
            pooled = true;
        }
    }
    auto finish = [&](int rc) {
        if (pooled) {
            cudaError_t e = cudaFreeAsync(scr, stream);   // stream-ordered: back to the pool after the last launch of this call
            if (e != cudaSuccess && !rc) rc = cuda_fail(e, "cudaFreeAsync(scratch)");
        }
        return rc;
    };
    if (need_gate) {   // layouts without a fused variant (odd group widths, misaligned rows): expand, then the plain kernels
        void *gbuf = scr + align256(need_long);
        if (int rc = spectre_gate_expand(gs->anchors, gs->bias, gs->eps, gs->pos, gs->pos_stride_b, gbuf, B, NG, gs->G, gs->Bk, F_half, stream_))
            return finish(rc);
        gate = gbuf;
        gs = nullptr;
    }
    // n_fft = 8192: the 4096-point kernel on (even rows, odd rows) tiles with the radix-2 combine in its middle pass
    if (kd)
        return finish(mix_dit(*st, *kd, v, v_dtype, v_stride_b, v_stride_n, gate, gs, mem, mem_stride, out, out_stride_b,
                              out_stride_n, B, n_io, n_fft, C, group_width, stream));
    // long transforms: one streaming radix-R pass, R interleaved 4096-point transforms in shared memory, one streaming pass
    if (ks)
        return finish(mix_two_pass(*st, *ks, v, v_dtype, v_stride_b, v_stride_n, gate, gs, mem, mem_stride, out, out_stride_b,
                                   out_stride_n, B, n_io, n_fft, C, group_width, reinterpret_cast<float *>(scr), stream));
    const float2 *tw = nullptr;
    if (int rc = get_twiddles(*st, *c.k, &tw)) return finish(rc);

    MixParams p;
    memset(&p, 0, sizeof(p));
    p.v = v;
    p.gate = reinterpret_cast<const float2 *>(gate);
    if (gs) p.gsrc = *gs;
    p.mem = reinterpret_cast<const float2 *>(mem);
    p.out = out;
    p.tw = tw;
    p.v_sb = v_stride_b;
    p.v_sn = v_stride_n;
    p.o_sb = out_stride_b;
    p.o_sn = out_stride_n;
    p.mem_stride = mem_stride;
    p.B = B;
    p.n_in = n_io;
    p.n_out = n_io;
    p.C = C;
    p.group_width = group_width;
    p.NG = NG;
    p.tiles_per_row = c.tiles_per_row;
    p.num_tiles = B * c.tiles_per_row;
    p.gate_tables = c.gate_tables;
    p.inv_n = 1.0f / (float)n_fft;
    p.prefetch = g_prefetch.load(std::memory_order_relaxed);
    p.timeline = g_timeline.load(std::memory_order_relaxed);
    p.skew_ns = g_skew_ns.load(std::memory_order_relaxed);
    p.sched = g_sched.load(std::memory_order_relaxed);
    p.sub_R = 1;
    p.sub_shift = 0;
    p.gw_shift = ilog2_exact(group_width);

    // TMA-fed variant when V's layout can be described to the TMA unit; otherwise direct 128-bit global loads
    alignas(64) CUtensorMap tmap, tmap_out;
    const int tile_ch = mode_channels(c.k->mode) * c.k->ncol;
    const bool use_tma = g_use_tma.load(std::memory_order_relaxed) != 0, use_tmem = g_use_tmem.load(std::memory_order_relaxed) != 0;
    const bool layout_ok = use_tma && c.k->tma_ok && tma_layout_ok(v, v_dtype, v_stride_b, v_stride_n) &&
                           tma_layout_ok(out, out_dtype, out_stride_b, out_stride_n);
    bool tmem = layout_ok && use_tmem && c.k->tmem_ok && (int)c.k->smem_bytes(c.gate_tables, true, true) <= st->max_smem_optin;
    bool tma = layout_ok && (tmem || (int)c.k->smem_bytes(c.gate_tables, true, false) <= st->max_smem_optin) &&
               make_v_tensor_map(&tmap, v, v_dtype, v_stride_b, v_stride_n, B, n_io, C,
                                 tmem ? c.k->tmem_box_rows : std::min(n_fft, spx::kTmaBoxRows), tile_ch,
                                 g_l2_promo.load(std::memory_order_relaxed)) &&
               make_v_tensor_map(&tmap_out, out, out_dtype, out_stride_b, out_stride_n, B, n_io, C,
                                 tmem ? c.k->tmem_box_rows : c.k->out_box_rows, tile_ch);
    if (!tma) tmem = false;
    // The narrow n_fft = 1024 variant (TMA landing in the working buffer, two CTAs per SM: what short launches such as BASELINE
    // configs[1] run) gains 1-2 % from the warp stagger the tensor-memory variants use, with a 500-cycle step
    // (profiles/r04w_ab_stagger_1024.txt: batch 8 / 16 / 24 / 32 / 40: +0.4 / +1.8 / +1.4 / +1.8 / -0.2 %).  Only with the knobs at
    // their defaults, so experiments keep full control.
    if (tma && !tmem && n_fft == 1024 && p.skew_ns == -350 && !(p.sched & 64)) {
        p.sched |= 64;
        p.skew_ns = -500;
    }
    const int occ = std::max(1, occupancy_of(*st, c, mem != nullptr, tma, tmem));
    // paired tile order (TMEM variant): the two channel tiles of one gate group back to back, gate row staged once for both
    p.pair_tiles = (tmem && (p.sched & 16) && c.gate_tables == 1 && c.tiles_per_row % 2 == 0 && group_width % (2 * tile_ch) == 0) ? 1 : 0;
    const int grid = std::min(p.pair_tiles ? p.num_tiles / 2 : p.num_tiles, st->sm_count * occ);
    cudaError_t e = c.k->launch(p, grid, mem != nullptr, tma ? &tmap : nullptr, tma ? &tmap_out : nullptr, tmem, stream);
    if (e != cudaSuccess) return finish(cuda_fail(e, "kernel launch"));
    return finish(0);
}
}  // namespace

int spectre_mix_fwd(const void *v, int v_dtype, int64_t v_stride_b, int64_t v_stride_n, const void *gate,
                    const void *mem, int64_t mem_stride, void *out, int out_dtype, int64_t out_stride_b,
                    int64_t out_stride_n, int B, int N, int n_fft, int C, int group_width, void *stream) {
    return mix_fwd_impl(v, v_dtype, v_stride_b, v_stride_n, gate, nullptr, mem, mem_stride, out, out_dtype, out_stride_b, out_stride_n,
                        B, N, n_fft, C, group_width, nullptr, 0, stream);
}

int spectre_mix_fwd_ws(const void *v, int v_dtype, int64_t v_stride_b, int64_t v_stride_n, const void *gate,
                       const void *mem, int64_t mem_stride, void *out, int out_dtype, int64_t out_stride_b,
                       int64_t out_stride_n, int B, int N, int n_fft, int C, int group_width, void *workspace,
                       size_t workspace_bytes, void *stream) {
    if (!workspace && workspace_bytes) return fail(SPECTRE_MIX_ERR_BAD_ARG, "workspace is null but workspace_bytes = %zu", workspace_bytes);
    return mix_fwd_impl(v, v_dtype, v_stride_b, v_stride_n, gate, nullptr, mem, mem_stride, out, out_dtype, out_stride_b, out_stride_n,
                        B, N, n_fft, C, group_width, workspace, workspace_bytes, stream);
}

int spectre_mix_fwd_anchors(const void *v, int v_dtype, int64_t v_stride_b, int64_t v_stride_n, const void *anchors,
                            const float *bias, const float *eps, const void *pos_phase, int64_t pos_stride_b, int G, int Bk,
                            const void *mem, int64_t mem_stride, void *out, int out_dtype, int64_t out_stride_b,
                            int64_t out_stride_n, int B, int N, int n_fft, int C, int group_width, void *workspace,
                            size_t workspace_bytes, void *stream) {
    if (!workspace && workspace_bytes) return fail(SPECTRE_MIX_ERR_BAD_ARG, "workspace is null but workspace_bytes = %zu", workspace_bytes);
    spx::GateSrc gs{reinterpret_cast<const float2 *>(anchors), bias, eps, reinterpret_cast<const float2 *>(pos_phase),
                    (long long)pos_stride_b, Bk, G};
    if (n_fft >= 2 && Bk >= 1 && B > 0) {
        if (int rc = spx::gate_interp_table(n_fft / 2 + 1, Bk, &gs.icoef, &gs.itap)) return rc;
    }
    return mix_fwd_impl(v, v_dtype, v_stride_b, v_stride_n, nullptr, &gs, mem, mem_stride, out, out_dtype, out_stride_b, out_stride_n,
                        B, N, n_fft, C, group_width, workspace, workspace_bytes, stream);
}

size_t spectre_mix_anchors_workspace_bytes(int v_dtype, int B, int N, int n_fft, int C, int group_width) {
    const size_t base = spectre_mix_workspace_bytes(v_dtype, B, N, n_fft, C, group_width);
    if (check_common(B, N, n_fft, C, group_width) || B == 0 || C == 0 || std::min(N, n_fft) == 0) return 0;
    // layouts the packed kernels cannot take (group width not a multiple of 4) materialise the gate next to the scratch
    const size_t gate = (group_width % 4 == 0) ? 0 : (size_t)B * (C / group_width) * (n_fft / 2 + 1) * sizeof(float2);
    return align256(base) + gate;
}

size_t spectre_mix_workspace_bytes(int v_dtype, int B, int N, int n_fft, int C, int group_width) {
    if (check_common(B, N, n_fft, C, group_width)) return 0;
    if (B == 0 || C == 0 || std::min(N, n_fft) == 0) return 0;
    // upper bound over layouts: the two-pass path is taken when the tensors allow the packed (QUAD) layout; a call that
    // falls back to a single-kernel variant ignores the workspace
    const int mode = (group_width % 4 == 0) ? spx::MODE_QUAD : ((group_width % 2 == 0) ? spx::MODE_PAIR : spx::MODE_REAL);
    if (!two_pass_kernel(n_fft, v_dtype, mode, group_width, nullptr, 0)) return 0;
    return (size_t)two_pass_rows(B, n_fft, C) * (size_t)n_fft * C * sizeof(float);
}

int spectre_mix_dgate(const void *v, const void *dy, int dtype, int64_t v_stride_b, int64_t v_stride_n, int64_t dy_stride_b,
                      int64_t dy_stride_n, void *dgate, int B, int N, int n_fft, int C, int group_width, void *stream_) {
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    if (int rc = check_common(B, N, n_fft, C, group_width)) return rc;
    if (dtype != SPECTRE_MIX_F32 && dtype != SPECTRE_MIX_BF16) return fail(SPECTRE_MIX_ERR_UNSUPPORTED, "dtype %d unsupported", dtype);
    const int n_io = std::min(N, n_fft);
    const int NG = C / group_width, F_half = n_fft / 2 + 1;
    if (B == 0 || C == 0) return 0;
    if (!dgate || ((!v || !dy) && n_io > 0)) return fail(SPECTRE_MIX_ERR_BAD_ARG, "null pointer (v=%p dy=%p dgate=%p)", v, dy, dgate);
    if (!aligned(dgate, 8)) return fail(SPECTRE_MIX_ERR_UNSUPPORTED, "dgate must be 8-byte aligned");
    DeviceState *st = nullptr;
    if (int rc = get_device_state(&st, nullptr)) return rc;
    // the fused variant: packed layout, TMEM-staged kernel of this n_fft, whole tiles inside one gate group
    const KernelEntry *k = nullptr;
    for (const KernelEntry &e : registry())
        if (e.n_fft == n_fft && e.io == dtype && e.mode == spx::MODE_QUAD && !e.sub && e.launch_dgate && e.tmem_ok) { k = &e; break; }
    const int tile_ch = k ? mode_channels(k->mode) * k->ncol : 0;
    const int mode = pick_mode(dtype, group_width, v, v_stride_b, v_stride_n, dy, dy_stride_b, dy_stride_n, nullptr, 0);
    if (!k || mode != spx::MODE_QUAD || group_width % tile_ch != 0 || !tma_layout_ok(v, dtype, v_stride_b, v_stride_n) ||
        !tma_layout_ok(dy, dtype, dy_stride_b, dy_stride_n) || (int)k->smem_bytes(1, true, true) > st->max_smem_optin)
        return fail(SPECTRE_MIX_ERR_UNSUPPORTED,
                    "no fused gate-gradient kernel for n_fft=%d dtype=%d group_width=%d (built for n_fft = 4096, group widths that are "
                    "multiples of 8, 16-byte aligned rows); use the two-spectrum formulation", n_fft, dtype, group_width);
    const float2 *tw = nullptr;
    if (int rc = get_twiddles(*st, *k, &tw)) return rc;
    cudaError_t e = cudaMemsetAsync(dgate, 0, (size_t)B * NG * F_half * sizeof(float2), stream);   // tiles of a group add into it
    if (e != cudaSuccess) return cuda_fail(e, "cudaMemsetAsync(dgate)");
    if (n_io == 0) return 0;
    MixParams p;
    memset(&p, 0, sizeof(p));
    p.v = v;
    p.out = dgate;
    p.tw = tw;
    p.v_sb = v_stride_b;
    p.v_sn = v_stride_n;
    p.B = B;
    p.n_in = n_io;
    p.n_out = n_io;
    p.C = C;
    p.group_width = group_width;
    p.NG = NG;
    p.tiles_per_row = (C / 4 + k->ncol - 1) / k->ncol;
    p.num_tiles = B * p.tiles_per_row;
    p.gate_tables = 1;
    p.inv_n = 1.0f / (float)n_fft;
    p.skew_ns = g_skew_ns.load(std::memory_order_relaxed);
    p.sched = g_sched.load(std::memory_order_relaxed) & ~(16 | 8 | 4);
    p.sub_R = 1;
    p.gw_shift = ilog2_exact(group_width);
    alignas(64) CUtensorMap tmap_v, tmap_dy;
    if (!make_v_tensor_map(&tmap_v, v, dtype, v_stride_b, v_stride_n, B, n_io, C, spx::kTmaBoxRows, tile_ch, 0) ||
        !make_v_tensor_map(&tmap_dy, dy, dtype, dy_stride_b, dy_stride_n, B, n_io, C, spx::kTmaBoxRows, tile_ch, 0))
        return fail(SPECTRE_MIX_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled failed for V / dY");
    const int grid = std::min(p.num_tiles, st->sm_count);
    e = k->launch_dgate(p, grid, &tmap_v, &tmap_dy, stream);
    if (e != cudaSuccess) return cuda_fail(e, "gate-gradient kernel launch");
    return 0;
}

int spectre_mix_plan(int v_dtype, int out_dtype, int has_mem, int B, int N, int n_fft, int C, int group_width,
                     spectre_mix_plan_info *info) {
    if (!info) return fail(SPECTRE_MIX_ERR_BAD_ARG, "info is null");
    if (int rc = check_common(B, N, n_fft, C, group_width)) return rc;
    if (v_dtype != out_dtype) return fail(SPECTRE_MIX_ERR_UNSUPPORTED, "dtype pair unsupported");
    DeviceState *st = nullptr;
    if (int rc = get_device_state(&st, nullptr)) return rc;
    const int mode = (group_width % 4 == 0) ? spx::MODE_QUAD : ((group_width % 2 == 0) ? spx::MODE_PAIR : spx::MODE_REAL);
    const KernelEntry *kd = (n_fft > 4096) ? dit_kernel(*st, n_fft, v_dtype, mode, std::min(N, n_fft), nullptr, 0) : nullptr;
    const KernelEntry *ks = kd ? nullptr : two_pass_kernel(n_fft, v_dtype, mode, group_width, nullptr, 0);
    const bool g_use_tma = ::g_use_tma.load(std::memory_order_relaxed) != 0, g_use_tmem = ::g_use_tmem.load(std::memory_order_relaxed) != 0;
    Choice c;
    if (kd) {
        c.k = kd;
        c.gate_tables = 2;
        c.tiles_per_row = C / 4;
    } else if (ks) {
        c.k = ks;
        c.gate_tables = 2;
        c.tiles_per_row = (C / 4 + ks->ncol - 1) / ks->ncol;
    } else if (int rc = choose(*st, n_fft, v_dtype, mode, C, group_width, &c, false, B)) return rc;
    memset(info, 0, sizeof(*info));
    info->n_fft = n_fft;
    for (int i = 0; i < 4; ++i) info->radix[i] = c.k->radix[i];
    info->tile_channels = kd ? 4 : mode_channels(c.k->mode) * c.k->ncol;
    info->dit = kd ? 2 : 1;
    info->threads = c.k->threads;
    info->ctas_per_sm = std::max(1, occupancy_of(*st, c, has_mem != 0, g_use_tma && c.k->tma_ok));
    info->smem_bytes = (int)c.k->smem_bytes(c.gate_tables, g_use_tma && c.k->tma_ok, g_use_tma && g_use_tmem && c.k->tmem_ok);
    info->grid = std::min(B * (ks ? n_fft / 4096 : 1) * c.tiles_per_row, st->sm_count * info->ctas_per_sm);
    info->launches = ks ? 3 : 1;   // long transforms: pre pass + 4096-point sub-transforms + post pass
    info->algorithmic_bytes = algorithmic_bytes(v_dtype, has_mem != 0, B, N, n_fft, C, group_width);
    info->workspace_bytes = (int64_t)spectre_mix_workspace_bytes(v_dtype, B, N, n_fft, C, group_width);
    return 0;
}

int spectre_rfft_fwd(const void *v, int v_dtype, int64_t v_stride_b, int64_t v_stride_n, void *spec, int B, int N,
                     int n_fft, int C, void *stream) {
    if (int rc = check_common(B, N, n_fft, C, 1)) return rc;
    if (v_dtype != SPECTRE_MIX_F32 && v_dtype != SPECTRE_MIX_BF16)
        return fail(SPECTRE_MIX_ERR_UNSUPPORTED, "dtype %d unsupported", v_dtype);
    if (B == 0 || C == 0) return 0;
    if (!spec || (!v && N > 0)) return fail(SPECTRE_MIX_ERR_BAD_ARG, "null pointer");
    if (!aligned(spec, 8)) return fail(SPECTRE_MIX_ERR_UNSUPPORTED, "spec must be 8-byte aligned");
    DeviceState *st = nullptr;
    if (int rc = get_device_state(&st, nullptr)) return rc;
    Choice c;
    if (int rc = choose(*st, n_fft, v_dtype, spx::MODE_REAL, C, 1, &c, /*no_gate=*/true)) return rc;
    if (!c.k->launch_rfft) return fail(SPECTRE_MIX_ERR_UNSUPPORTED, "internal: no rfft variant");
    const float2 *tw = nullptr;
    if (int rc = get_twiddles(*st, *c.k, &tw)) return rc;
    MixParams p;
    memset(&p, 0, sizeof(p));
    p.v = v;
    p.out = spec;
    p.tw = tw;
    p.v_sb = v_stride_b;
    p.v_sn = v_stride_n;
    p.o_sb = (long long)(n_fft / 2 + 1) * C;  // complex elements
    p.o_sn = C;
    p.B = B;
    p.n_in = std::min(N, n_fft);
    p.n_out = 0;
    p.C = C;
    p.group_width = 1;
    p.NG = C;
    p.tiles_per_row = c.tiles_per_row;
    p.num_tiles = B * c.tiles_per_row;
    p.gate_tables = 0;
    p.inv_n = 1.0f;
    p.prefetch = 0;
    p.sub_R = 1;
    p.gw_shift = 0;
    const int grid = std::min(p.num_tiles, st->sm_count * std::max(1, c.k->minb));
    cudaError_t e = c.k->launch_rfft(p, grid, reinterpret_cast<cudaStream_t>(stream));
    if (e != cudaSuccess) return cuda_fail(e, "rfft kernel launch");
    return 0;
}

// ---------------------------------------------------------------- host-buffer entry point
namespace {
constexpr int kHostStreams = 4;   // chunks in flight: H2D of one, kernel of another and D2H of a third overlap, one spare
struct HostCtx {
    cudaStream_t s[kHostStreams] = {};
    void *dv[kHostStreams] = {}, *dg[kHostStreams] = {}, *dout[kHostStreams] = {};
    void *dws[kHostStreams] = {};   // long-context workspace, one per stream: chunks in flight never share an intermediate
    size_t cap_v = 0, cap_g = 0, cap_ws = 0;
    void *dmem = nullptr;
    size_t cap_mem = 0;
};
std::mutex g_host_mu;
std::map<int, HostCtx> g_host;
// Chunk schedule of the host entry.  The call is PCIe-bound (both copy engines busy for ~40 ms at the metric shape), so what
// matters is (a) a short pipeline fill before both engines run and a short drain after -- small chunks at the two ends -- and
// (b) few copy -> kernel -> copy hand-overs in between -- large chunks in the middle.  Rows per chunk ramp up by doubling from
// ~12 MB of V to ~100 MB and back down (1, 2, 4, 8, ..., 8, 4, 2, 1 rows at seq 4096 x d 768).  Measured on the metric
// shape (148 rows, profiles/r04b_e2e_probe.txt, r04d_e2e_ab.txt): uniform 32 MB chunks 44.1 ms, uniform 64 MB 42.4 ms, ramped
// 42.2 ms; the same bytes as two monolithic copies 39.6 ms.  SPECTRE_MIX_HOST_CHUNK_MB (experiment knob) forces uniform chunks
// of that size, SPECTRE_MIX_HOST_CHUNK_MAX_MB moves the top of the ramp.
std::vector<int> host_chunk_schedule(int B, size_t row_v) {
    size_t lo_bytes = 12u << 20, hi_bytes = 100u << 20;
    bool uniform = false;
    if (const char *env = getenv("SPECTRE_MIX_HOST_CHUNK_MB")) {
        const long mb = strtol(env, nullptr, 10);
        if (mb > 0 && mb <= 4096) { lo_bytes = hi_bytes = (size_t)mb << 20; uniform = true; }
    }
    if (const char *env = getenv("SPECTRE_MIX_HOST_CHUNK_MAX_MB")) {
        const long mb = strtol(env, nullptr, 10);
        if (mb > 0 && mb <= 4096) hi_bytes = std::max(lo_bytes, (size_t)mb << 20);
    }
    auto rows_for = [&](size_t bytes) { return (int)std::max<size_t>(1, std::min<size_t>((size_t)B, bytes / std::max<size_t>(row_v, 1))); };
    const int r_lo = rows_for(lo_bytes), r_hi = std::max(r_lo, rows_for(hi_bytes));
    std::vector<int> ramp, sched;
    long long ramp_sum = 0;
    if (!uniform)
        for (long long r = r_lo; r < r_hi && 2 * (ramp_sum + r) + r_hi <= B; r *= 2) { ramp.push_back((int)r); ramp_sum += r; }
    int left = B - (int)(2 * ramp_sum);
    sched = ramp;
    const int top = ramp.empty() ? (uniform ? r_hi : r_lo) : r_hi;   // too few rows for a ramp: the small chunk size throughout
    while (left > 0) { const int r = std::min(top, left); sched.push_back(r); left -= r; }
    for (size_t i = ramp.size(); i-- > 0;) sched.push_back(ramp[i]);
    return sched;
}
}  // namespace

int spectre_mix_fwd_host(const float *v, const float *gate, const float *mem, float *out, int B, int N, int n_fft,
                         int C, int group_width) {
    if (int rc = check_common(B, N, n_fft, C, group_width)) return rc;
    const int n_io = std::min(N, n_fft);
    if (B == 0 || C == 0 || n_io == 0) return 0;
    if (!v || !gate || !out) return fail(SPECTRE_MIX_ERR_BAD_ARG, "null pointer");
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(SPECTRE_MIX_ERR_NO_DEVICE, "no CUDA device: %s", cudaGetErrorString(e));
    }
    std::lock_guard<std::mutex> lock(g_host_mu);
    HostCtx &h = g_host[dev];
    const size_t fh = n_fft / 2 + 1, ng = C / group_width;
    const size_t row_v = (size_t)N * C * 4, row_o = (size_t)n_io * C * 4, row_g = ng * fh * 8;
    const std::vector<int> sched = host_chunk_schedule(B, row_v);   // rows of every chunk, in order
    int rows = 1;   // largest chunk: sizes the per-stream device buffers
    for (int r : sched) rows = std::max(rows, r);
    const size_t need_v = std::max(row_v, row_o) * rows, need_g = row_g * rows;
    for (int i = 0; i < kHostStreams; ++i) {
        if (!h.s[i] && (e = cudaStreamCreateWithFlags(&h.s[i], cudaStreamNonBlocking)) != cudaSuccess)
            return cuda_fail(e, "cudaStreamCreate");
    }
    if (need_v > h.cap_v) {
        for (int i = 0; i < kHostStreams; ++i) {
            cudaFree(h.dv[i]);
            cudaFree(h.dout[i]);
            if ((e = cudaMalloc(&h.dv[i], need_v)) != cudaSuccess) return cuda_fail(e, "cudaMalloc(V chunk)");
            if ((e = cudaMalloc(&h.dout[i], need_v)) != cudaSuccess) return cuda_fail(e, "cudaMalloc(out chunk)");
        }
        h.cap_v = need_v;
    }
    if (need_g > h.cap_g) {
        for (int i = 0; i < kHostStreams; ++i) {
            cudaFree(h.dg[i]);
            if ((e = cudaMalloc(&h.dg[i], need_g)) != cudaSuccess) return cuda_fail(e, "cudaMalloc(gate chunk)");
        }
        h.cap_g = need_g;
    }
    const size_t need_ws = spectre_mix_workspace_bytes(SPECTRE_MIX_F32, rows, N, n_fft, C, group_width);
    if (need_ws > h.cap_ws) {
        for (int i = 0; i < kHostStreams; ++i) {
            cudaFree(h.dws[i]);   // every stream was synchronised when the previous call returned
            h.dws[i] = nullptr;
            if ((e = cudaMalloc(&h.dws[i], need_ws)) != cudaSuccess) { h.cap_ws = 0; return cuda_fail(e, "cudaMalloc(long-context workspace)"); }
        }
        h.cap_ws = need_ws;
    }
    if (mem) {
        const size_t need_m = fh * (size_t)C * 8;
        if (need_m > h.cap_mem) {
            cudaFree(h.dmem);
            if ((e = cudaMalloc(&h.dmem, need_m)) != cudaSuccess) return cuda_fail(e, "cudaMalloc(memory)");
            h.cap_mem = need_m;
        }
        if ((e = cudaMemcpyAsync(h.dmem, mem, need_m, cudaMemcpyHostToDevice, h.s[0])) != cudaSuccess)
            return cuda_fail(e, "cudaMemcpyAsync(memory)");
        if ((e = cudaStreamSynchronize(h.s[0])) != cudaSuccess) return cuda_fail(e, "sync(memory)");
    }
    int b0 = 0;
    for (size_t chunk = 0; chunk < sched.size(); b0 += sched[chunk], ++chunk) {
        const int nb = sched[chunk];
        const int i = (int)(chunk % kHostStreams);
        cudaStream_t s = h.s[i];
        if ((e = cudaMemcpyAsync(h.dv[i], v + (size_t)b0 * N * C, row_v * nb, cudaMemcpyHostToDevice, s)) != cudaSuccess)
            return cuda_fail(e, "H2D V");
        if ((e = cudaMemcpyAsync(h.dg[i], reinterpret_cast<const char *>(gate) + (size_t)b0 * row_g, row_g * nb,
                                 cudaMemcpyHostToDevice, s)) != cudaSuccess)
            return cuda_fail(e, "H2D gate");
        int rc = spectre_mix_fwd_ws(h.dv[i], SPECTRE_MIX_F32, (int64_t)N * C, C, h.dg[i], mem ? h.dmem : nullptr, C, h.dout[i],
                                    SPECTRE_MIX_F32, (int64_t)n_io * C, C, nb, N, n_fft, C, group_width,
                                    need_ws ? h.dws[i] : nullptr, need_ws ? h.cap_ws : 0, s);
        if (rc) return rc;
        if ((e = cudaMemcpyAsync(out + (size_t)b0 * n_io * C, h.dout[i], row_o * nb, cudaMemcpyDeviceToHost, s)) !=
            cudaSuccess)
            return cuda_fail(e, "D2H out");
    }
    for (int i = 0; i < kHostStreams; ++i)
        if ((e = cudaStreamSynchronize(h.s[i])) != cudaSuccess) return cuda_fail(e, "stream sync");
    return 0;
}

int spectre_mix_host_schedule(int B, int N, int C, int *rows_out, int cap) {
    if (B < 0 || N <= 0 || C <= 0 || (!rows_out && cap > 0)) return -1;
    if (B == 0) return 0;
    const std::vector<int> sched = host_chunk_schedule(B, (size_t)N * C * 4);
    for (int i = 0; i < (int)sched.size() && i < cap; ++i) rows_out[i] = sched[i];
    return (int)sched.size();
}

int spectre_mix_host_release(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        return 0;   // no device: nothing was ever allocated
    }
    std::lock_guard<std::mutex> lock(g_host_mu);
    auto it = g_host.find(dev);
    if (it == g_host.end()) return 0;
    HostCtx &h = it->second;
    for (int i = 0; i < kHostStreams; ++i) {
        if (h.s[i]) cudaStreamSynchronize(h.s[i]);
        cudaFree(h.dv[i]);
        cudaFree(h.dout[i]);
        cudaFree(h.dg[i]);
        cudaFree(h.dws[i]);
        if (h.s[i]) cudaStreamDestroy(h.s[i]);
    }
    cudaFree(h.dmem);
    g_host.erase(it);
    cudaGetLastError();
    return 0;
}

}  // extern "C"
