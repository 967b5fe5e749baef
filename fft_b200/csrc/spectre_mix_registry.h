// Internal: table of compiled kernel variants, filled by the per-size instantiation files.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <string.h>

#include "spectre_mix_kernel.cuh"

namespace spx {

struct KernelEntry {
    int n_fft;
    int radix[4];
    int mode;      // Mode
    int io;        // SPECTRE_MIX_F32 / SPECTRE_MIX_BF16 (V and out share it)
    int ncol;      // elements (CH channels each) per tile row
    int threads;
    int minb;
    int twn;       // twiddle table entries
    size_t (*smem_bytes)(int gate_tables, bool tma, bool tmem);
    int out_box_rows;   // rows per TMA store box (TMA variant)
    int tmem_box_rows;  // rows per TMA box, loads and stores, of the TMEM-staged variant
    // tmap != nullptr selects the TMA-fed variant (only when tma_ok)
    // p.gate == nullptr selects the in-kernel gate generator (p.gsrc = anchors; only when anch_ok)
    cudaError_t (*launch)(const MixParams &p, int grid, bool has_mem, const CUtensorMap *tmap_in, const CUtensorMap *tmap_out,
                          bool tmem, cudaStream_t st);
    int (*occupancy)(int gate_tables, bool has_mem, bool tma, bool tmem);
    int tma_ok;    // 1: a TMA-fed variant exists (packed mode, landed row >= 16 bytes)
    int tmem_ok;   // 1: a TMEM-staged variant exists (tile I/O parked in tensor memory by a helper warpgroup)
    int sub;       // 1: sub-transform variant of the long-context two-pass path (complex input, strided gate gather)
    int anch_ok;   // 1: variants that evaluate the gate from anchors inside the gate staging exist (packed mode)
    int dit;       // 2: DIT2 variant -- the two tile columns are the even / odd rows of a transform of length 2 * n_fft
    // forward half only (half spectrum out); MODE_REAL variants only, else nullptr
    cudaError_t (*launch_rfft)(const MixParams &p, int grid, cudaStream_t st);
    // gate gradient: TMEM-staged packed variants only, else nullptr; tmap_v / tmap_dy describe V and dY
    cudaError_t (*launch_dgate)(const MixParams &p, int grid, const CUtensorMap *tmap_v, const CUtensorMap *tmap_dy, cudaStream_t st);
};

template <class PL, int MODE, int NCOL, int NT, int MINB, class TIO>
struct Launcher {
    using SM = Smem<PL, MODE, NCOL>;
    static size_t smem_bytes(int gate_tables, bool tma, bool tmem = false) {
        return SM::bytes(gate_tables, tma && kTma, sizeof(typename Lin<MODE, TIO>::T) * NCOL, tma && tmem && kTmem);
    }
    // TMA delivers rows of NCOL elements; the box's inner extent must be a multiple of 16 bytes
    static constexpr bool kTma = (MODE == MODE_QUAD) && ((sizeof(TIO) * 4 * NCOL) % 16 == 0);
    // TMEM staging: packed tiles (fp32 or bf16 in HBM, fp32 in tensor memory), one stage-0 butterfly per thread, 1 CTA per SM, two tiles fit the 512 columns
    static constexpr bool kTmem = kTma && NT >= kSepProducerMinThreads && NT == NCOL * PL::L(0) &&
                                  (PL::L(0) * NCOL) % 128 == 0 && 128 % NCOL == 0 && MINB == 1 && (PL::N * NCOL * 4 / 128 * 2 <= 512);
    // in-kernel gate generation (spectre_mix_fwd_anchors): built for the packed mode, the layout of every Spectre model with
    // 4 | group_width; other layouts take the materialised-gate kernels behind spectre_gate_expand
    static constexpr bool kAnch = (MODE == MODE_QUAD);
    template <bool HAS_MEM, bool TMA, bool TMEM, bool ANCH = false>
    static const void *fn() {
        return reinterpret_cast<const void *>(
            &spectre_mix_kernel<PL, MODE, NCOL, NT, MINB, TIO, TIO, HAS_MEM, false, (TMA && kTma), (TMA && TMEM && kTmem), (ANCH && kAnch)>);
    }
    static const void *pick(bool has_mem, bool tma, bool tmem = false, bool anch = false) {
        if constexpr (PL::kDit) {   // DIT2: TMEM-staged variants only
            if (anch) return has_mem ? fn<true, true, true, true>() : fn<false, true, true, true>();
            return has_mem ? fn<true, true, true>() : fn<false, true, true>();
        } else {
            if (anch && kAnch) {
                if (tma && tmem && kTmem) return has_mem ? fn<true, true, true, true>() : fn<false, true, true, true>();
                if (tma && kTma) return has_mem ? fn<true, true, false, true>() : fn<false, true, false, true>();
                return has_mem ? fn<true, false, false, true>() : fn<false, false, false, true>();
            }
            if (tma && tmem && kTmem) return has_mem ? fn<true, true, true>() : fn<false, true, true>();
            if (tma && kTma) return has_mem ? fn<true, true, false>() : fn<false, true, false>();
            return has_mem ? fn<true, false, false>() : fn<false, false, false>();
        }
    }
    static cudaError_t launch(const MixParams &p, int grid, bool has_mem, const CUtensorMap *tmap, const CUtensorMap *tmap_out,
                              bool tmem, cudaStream_t st) {
        const size_t sm = smem_bytes(p.gate_tables, tmap != nullptr, tmem);
        const void *f = pick(has_mem, tmap != nullptr, tmem, p.gate == nullptr);
        // opt in to > 48 KB dynamic shared memory (cheap; the driver caches the attribute per function)
        cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
        MixParams pc = p;
        alignas(64) CUtensorMap tm, tmo;
        if (tmap) { tm = *tmap; tmo = *tmap_out; } else { memset(&tm, 0, sizeof(tm)); memset(&tmo, 0, sizeof(tmo)); }
        void *args[] = {&pc, &tm, &tmo};
        // Programmatic dependent launch (sched bit 7 switches it OFF): the kernel's set-up may overlap the tail of the previous
        // kernel of the stream; it waits with griddepcontrol.wait before it touches any tensor (griddep_wait in the kernel header)
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(NT + ((tmap != nullptr && kTma && NT >= kSepProducerMinThreads) ? kProducerThreads : 0));
        cfg.dynamicSmemBytes = sm;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = (p.sched & 128) ? 0 : 1;
        return cudaLaunchKernelExC(&cfg, f, args);
    }
    static cudaError_t launch_rfft(const MixParams &p, int grid, cudaStream_t st) {
        static_assert(MODE == MODE_REAL, "rfft-only is built for MODE_REAL");
        const size_t sm = smem_bytes(0, false, false);
        const void *f = reinterpret_cast<const void *>(&spectre_mix_kernel<PL, MODE, NCOL, NT, MINB, TIO, TIO, false, true>);
        cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
        MixParams pc = p;
        alignas(64) CUtensorMap tm, tmo;
        memset(&tm, 0, sizeof(tm));
        memset(&tmo, 0, sizeof(tmo));
        void *args[] = {&pc, &tm, &tmo};
        return cudaLaunchKernel(f, dim3(grid), dim3(NT), args, sm, st);
    }
    static cudaError_t launch_dgate(const MixParams &p, int grid, const CUtensorMap *tmap_v, const CUtensorMap *tmap_dy, cudaStream_t st) {
        if constexpr (kTmem && PL::NS == 3 && !PL::kSub && !PL::kDit && PL::R(2) == 16) {
            const size_t sm = smem_bytes(1, true, true);
            const void *f = reinterpret_cast<const void *>(
                &spectre_mix_kernel<PL, MODE, NCOL, NT, MINB, TIO, TIO, false, false, true, true, false, true>);
            cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
            if (e != cudaSuccess) return e;
            MixParams pc = p;
            alignas(64) CUtensorMap tm = *tmap_v, tmo = *tmap_dy;
            void *args[] = {&pc, &tm, &tmo};
            return cudaLaunchKernel(f, dim3(grid), dim3(NT + kProducerThreads), args, sm, st);
        } else {
            return cudaErrorNotSupported;
        }
    }
    static constexpr bool kDgate = kTmem && PL::NS == 3 && !PL::kSub && !PL::kDit && PL::R(2) == 16;
    static int occupancy(int gate_tables, bool has_mem, bool tma, bool tmem) {
        const size_t sm = smem_bytes(gate_tables, tma, tmem);
        const void *f = pick(has_mem, tma, tmem);
        if (cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, f, NT + ((tma && kTma && NT >= kSepProducerMinThreads) ? kProducerThreads : 0), sm) != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
        return nb;
    }
};

template <class PL, int MODE, int NCOL, int NT, int MINB, class TIO>
struct RfftPtr {
    static constexpr decltype(&Launcher<PL, MODE_REAL, NCOL, NT, MINB, TIO>::launch_rfft) get() { return nullptr; }
};
template <class PL, int NCOL, int NT, int MINB, class TIO>
struct RfftPtr<PL, MODE_REAL, NCOL, NT, MINB, TIO> {
    static constexpr decltype(&Launcher<PL, MODE_REAL, NCOL, NT, MINB, TIO>::launch_rfft) get() {
        return &Launcher<PL, MODE_REAL, NCOL, NT, MINB, TIO>::launch_rfft;
    }
};

#define SPX_ENTRY(R0, R1, R2, R3, MODE, NCOL, NT, MINB, TIO, IOCODE)                                         \
    {                                                                                                         \
        (R0) * (R1) * (R2) * (R3), {R0, R1, R2, R3}, MODE, IOCODE, NCOL, NT, MINB,                            \
            ::spx::Plan<R0, R1, R2, R3>::TWN,                                                                 \
            &::spx::Launcher<::spx::Plan<R0, R1, R2, R3>, MODE, NCOL, NT, MINB, TIO>::smem_bytes,             \
            ::spx::Smem<::spx::Plan<R0, R1, R2, R3>, MODE, NCOL>::OUT_BOX_ROWS, ::spx::tmem_box_rows(NCOL),   \
            &::spx::Launcher<::spx::Plan<R0, R1, R2, R3>, MODE, NCOL, NT, MINB, TIO>::launch,                 \
            &::spx::Launcher<::spx::Plan<R0, R1, R2, R3>, MODE, NCOL, NT, MINB, TIO>::occupancy,              \
            ::spx::Launcher<::spx::Plan<R0, R1, R2, R3>, MODE, NCOL, NT, MINB, TIO>::kTma ? 1 : 0,            \
            ::spx::Launcher<::spx::Plan<R0, R1, R2, R3>, MODE, NCOL, NT, MINB, TIO>::kTmem ? 1 : 0, 0,        \
            ::spx::Launcher<::spx::Plan<R0, R1, R2, R3>, MODE, NCOL, NT, MINB, TIO>::kAnch ? 1 : 0, 0,        \
            ::spx::RfftPtr<::spx::Plan<R0, R1, R2, R3>, MODE, NCOL, NT, MINB, TIO>::get(),                    \
            ::spx::Launcher<::spx::Plan<R0, R1, R2, R3>, MODE, NCOL, NT, MINB, TIO>::kDgate                   \
                ? &::spx::Launcher<::spx::Plan<R0, R1, R2, R3>, MODE, NCOL, NT, MINB, TIO>::launch_dgate : nullptr \
    }

#define SPX_ENTRY_SUB(R0, R1, R2, R3, MODE, NCOL, NT, MINB, TIO, IOCODE)                                     \
    {                                                                                                         \
        (R0) * (R1) * (R2) * (R3), {R0, R1, R2, R3}, MODE, IOCODE, NCOL, NT, MINB,                            \
            ::spx::Plan<R0, R1, R2, R3, true>::TWN,                                                           \
            &::spx::Launcher<::spx::Plan<R0, R1, R2, R3, true>, MODE, NCOL, NT, MINB, TIO>::smem_bytes,       \
            ::spx::Smem<::spx::Plan<R0, R1, R2, R3, true>, MODE, NCOL>::OUT_BOX_ROWS, ::spx::tmem_box_rows(NCOL), \
            &::spx::Launcher<::spx::Plan<R0, R1, R2, R3, true>, MODE, NCOL, NT, MINB, TIO>::launch,           \
            &::spx::Launcher<::spx::Plan<R0, R1, R2, R3, true>, MODE, NCOL, NT, MINB, TIO>::occupancy,        \
            ::spx::Launcher<::spx::Plan<R0, R1, R2, R3, true>, MODE, NCOL, NT, MINB, TIO>::kTma ? 1 : 0,      \
            ::spx::Launcher<::spx::Plan<R0, R1, R2, R3, true>, MODE, NCOL, NT, MINB, TIO>::kTmem ? 1 : 0, 1,  \
            ::spx::Launcher<::spx::Plan<R0, R1, R2, R3, true>, MODE, NCOL, NT, MINB, TIO>::kAnch ? 1 : 0, 0,  \
            nullptr, nullptr                                                                                           \
    }

#define SPX_ENTRY_DIT(R0, R1, R2, R3, MODE, NCOL, NT, MINB, TIO, IOCODE)                                     \
    {                                                                                                         \
        (R0) * (R1) * (R2) * (R3), {R0, R1, R2, R3}, MODE, IOCODE, NCOL, NT, MINB,                            \
            ::spx::Plan<R0, R1, R2, R3, 2>::TWN,                                                              \
            &::spx::Launcher<::spx::Plan<R0, R1, R2, R3, 2>, MODE, NCOL, NT, MINB, TIO>::smem_bytes,          \
            ::spx::Smem<::spx::Plan<R0, R1, R2, R3, 2>, MODE, NCOL>::OUT_BOX_ROWS, ::spx::tmem_box_rows(NCOL), \
            &::spx::Launcher<::spx::Plan<R0, R1, R2, R3, 2>, MODE, NCOL, NT, MINB, TIO>::launch,              \
            &::spx::Launcher<::spx::Plan<R0, R1, R2, R3, 2>, MODE, NCOL, NT, MINB, TIO>::occupancy,           \
            ::spx::Launcher<::spx::Plan<R0, R1, R2, R3, 2>, MODE, NCOL, NT, MINB, TIO>::kTma ? 1 : 0,         \
            ::spx::Launcher<::spx::Plan<R0, R1, R2, R3, 2>, MODE, NCOL, NT, MINB, TIO>::kTmem ? 1 : 0, 0,     \
            ::spx::Launcher<::spx::Plan<R0, R1, R2, R3, 2>, MODE, NCOL, NT, MINB, TIO>::kAnch ? 1 : 0, 2,     \
            nullptr, nullptr                                                                                  \
    }

// one table per instantiation file
const KernelEntry *table_small(int *count);    // n_fft 32 .. 512
const KernelEntry *table_1024(int *count);
const KernelEntry *table_2048(int *count);
const KernelEntry *table_4096(int *count);
const KernelEntry *table_8192(int *count);
const KernelEntry *table_16384(int *count);

}  // namespace spx
