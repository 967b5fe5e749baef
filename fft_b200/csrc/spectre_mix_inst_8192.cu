// Kernel variants for n_fft group "8192": one table per file so that the instantiations compile in parallel (make -j).
// Registry order matters: the first entry of a (n_fft, dtype, mode) that fits is the default (spectre_mix_api.cu: choose()).
#include "spectre_mix_registry.h"
#include "../../include/spectre_mix.h"

namespace spx {
namespace {
const KernelEntry kTable[] = {
    // DIT2 variant (default at n_fft = 8192 fp32 when the row count is even): the 4096-point TMEM-staged kernel on tiles whose two
    // columns are the even / odd rows of 4 channels, the radix-2 combine in the middle pass (registers + 64 shuffles per thread)
    SPX_ENTRY_DIT(16, 16, 16, 1, MODE_QUAD, 2, 512, 1, float, SPECTRE_MIX_F32),
    // single-kernel TMEM-staged variant: radix-16 first (512 stage-0 butterflies = one per thread, rows u + 512 m in one tensor-
    // memory lane), 4-channel tiles (16-byte rows, 128 KB per tile), the radix-2 stage as an extra in-place pass
    SPX_ENTRY(16, 2, 16, 16, MODE_QUAD, 1, 512, 1, float, SPECTRE_MIX_F32),
    SPX_ENTRY(2, 16, 16, 16, MODE_QUAD, 1, 512, 1, float, SPECTRE_MIX_F32),
    SPX_ENTRY(2, 16, 16, 16, MODE_QUAD, 1, 512, 1, __nv_bfloat16, SPECTRE_MIX_BF16),
    SPX_ENTRY(2, 16, 16, 16, MODE_PAIR, 2, 512, 1, float, SPECTRE_MIX_F32),
    SPX_ENTRY(2, 16, 16, 16, MODE_PAIR, 2, 512, 1, __nv_bfloat16, SPECTRE_MIX_BF16),
    SPX_ENTRY(2, 16, 16, 16, MODE_REAL, 2, 512, 1, float, SPECTRE_MIX_F32),
    SPX_ENTRY(2, 16, 16, 16, MODE_REAL, 2, 512, 1, __nv_bfloat16, SPECTRE_MIX_BF16),
    SPX_ENTRY(2, 16, 16, 16, MODE_PAIR, 1, 512, 1, float, SPECTRE_MIX_F32),   // narrow fallback: one gate group per tile
    SPX_ENTRY(2, 16, 16, 16, MODE_PAIR, 1, 512, 1, __nv_bfloat16, SPECTRE_MIX_BF16),   // narrow fallback: one gate group per tile
    SPX_ENTRY(2, 16, 16, 16, MODE_REAL, 1, 512, 1, float, SPECTRE_MIX_F32),   // narrow fallback: one gate group per tile
    SPX_ENTRY(2, 16, 16, 16, MODE_REAL, 1, 512, 1, __nv_bfloat16, SPECTRE_MIX_BF16),   // narrow fallback: one gate group per tile
};
}  // namespace
const KernelEntry *table_8192(int *count) {
    *count = (int)(sizeof(kTable) / sizeof(kTable[0]));
    return kTable;
}
}  // namespace spx
