// Internal: table of compiled kernel variants, filled by the per-size instantiation files.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

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
    size_t (*smem_bytes)(int gate_tables);
    cudaError_t (*launch)(const MixParams &p, int grid, bool has_mem, cudaStream_t st);
    int (*occupancy)(int gate_tables, bool has_mem);
    // forward half only (half spectrum out); MODE_REAL variants only, else nullptr
    cudaError_t (*launch_rfft)(const MixParams &p, int grid, cudaStream_t st);
};

template <class PL, int MODE, int NCOL, int NT, int MINB, class TIO>
struct Launcher {
    using SM = Smem<PL, MODE, NCOL>;
    static size_t smem_bytes(int gate_tables) { return SM::bytes(gate_tables); }
    template <bool HAS_MEM>
    static const void *fn() {
        return reinterpret_cast<const void *>(&spectre_mix_kernel<PL, MODE, NCOL, NT, MINB, TIO, TIO, HAS_MEM>);
    }
    static cudaError_t launch(const MixParams &p, int grid, bool has_mem, cudaStream_t st) {
        const size_t sm = smem_bytes(p.gate_tables);
        const void *f = has_mem ? fn<true>() : fn<false>();
        // opt in to > 48 KB dynamic shared memory (cheap; the driver caches the attribute per function)
        cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
        MixParams pc = p;
        void *args[] = {&pc};
        return cudaLaunchKernel(f, dim3(grid), dim3(NT), args, sm, st);
    }
    static cudaError_t launch_rfft(const MixParams &p, int grid, cudaStream_t st) {
        static_assert(MODE == MODE_REAL, "rfft-only is built for MODE_REAL");
        const size_t sm = smem_bytes(0);
        const void *f = reinterpret_cast<const void *>(&spectre_mix_kernel<PL, MODE, NCOL, NT, MINB, TIO, TIO, false, true>);
        cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
        MixParams pc = p;
        void *args[] = {&pc};
        return cudaLaunchKernel(f, dim3(grid), dim3(NT), args, sm, st);
    }
    static int occupancy(int gate_tables, bool has_mem) {
        const size_t sm = smem_bytes(gate_tables);
        const void *f = has_mem ? fn<true>() : fn<false>();
        if (cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess) {
            cudaGetLastError();
            return 0;
        }
        int nb = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, f, NT, sm) != cudaSuccess) {
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
            &::spx::Launcher<::spx::Plan<R0, R1, R2, R3>, MODE, NCOL, NT, MINB, TIO>::launch,                 \
            &::spx::Launcher<::spx::Plan<R0, R1, R2, R3>, MODE, NCOL, NT, MINB, TIO>::occupancy,              \
            ::spx::RfftPtr<::spx::Plan<R0, R1, R2, R3>, MODE, NCOL, NT, MINB, TIO>::get()                     \
    }

// one table per instantiation file
const KernelEntry *table_small(int *count);    // n_fft 32 .. 512
const KernelEntry *table_1024(int *count);
const KernelEntry *table_2048(int *count);
const KernelEntry *table_4096(int *count);
const KernelEntry *table_8192(int *count);
const KernelEntry *table_16384(int *count);

}  // namespace spx
