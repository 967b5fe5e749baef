// Kernel variants for n_fft group "1024": one table per file so that the instantiations compile in parallel (make -j).
// Registry order matters: the first entry of a (n_fft, dtype, mode) that fits is the default (spectre_mix_api.cu: choose()).
#include "spectre_mix_registry.h"
#include "../../include/spectre_mix.h"

namespace spx {
namespace {
const KernelEntry kTable[] = {
    // wide-row TMEM-staged variant (default): 1024 = 16 x 16 x 4, 32-channel tiles = 128-byte rows, 1024 x 32 x 4 B = the same
    // 128 KB tile, tensor-memory staging and helper warpgroup as the 4096 kernel
    SPX_ENTRY(16, 16, 4, 1, MODE_QUAD, 8, 512, 1, float, SPECTRE_MIX_F32),
    SPX_ENTRY(16, 16, 4, 1, MODE_QUAD, 8, 512, 1, __nv_bfloat16, SPECTRE_MIX_BF16),
    // 16-channel tiles first (default): 64-byte fp32 / 32-byte bf16 rows (measured 4160 vs 3950 GB/s fp32 at B = 256)
    SPX_ENTRY(4, 16, 16, 1, MODE_QUAD, 4, 256, 2, float, SPECTRE_MIX_F32),
    SPX_ENTRY(4, 16, 16, 1, MODE_QUAD, 4, 256, 2, __nv_bfloat16, SPECTRE_MIX_BF16),
    SPX_ENTRY(4, 16, 16, 1, MODE_QUAD, 2, 128, 4, float, SPECTRE_MIX_F32),
    SPX_ENTRY(4, 16, 16, 1, MODE_QUAD, 2, 128, 4, __nv_bfloat16, SPECTRE_MIX_BF16),
    SPX_ENTRY(4, 16, 16, 1, MODE_PAIR, 4, 256, 2, float, SPECTRE_MIX_F32),
    SPX_ENTRY(4, 16, 16, 1, MODE_PAIR, 4, 256, 2, __nv_bfloat16, SPECTRE_MIX_BF16),
    SPX_ENTRY(4, 16, 16, 1, MODE_REAL, 4, 256, 2, float, SPECTRE_MIX_F32),
    SPX_ENTRY(4, 16, 16, 1, MODE_REAL, 4, 256, 2, __nv_bfloat16, SPECTRE_MIX_BF16),
    SPX_ENTRY(4, 16, 16, 1, MODE_PAIR, 1, 64, 2, float, SPECTRE_MIX_F32),   // narrow fallback: one gate group per tile
    SPX_ENTRY(4, 16, 16, 1, MODE_PAIR, 1, 64, 2, __nv_bfloat16, SPECTRE_MIX_BF16),   // narrow fallback: one gate group per tile
    SPX_ENTRY(4, 16, 16, 1, MODE_REAL, 1, 64, 2, float, SPECTRE_MIX_F32),   // narrow fallback: one gate group per tile
    SPX_ENTRY(4, 16, 16, 1, MODE_REAL, 1, 64, 2, __nv_bfloat16, SPECTRE_MIX_BF16),   // narrow fallback: one gate group per tile
};
}  // namespace
const KernelEntry *table_1024(int *count) {
    *count = (int)(sizeof(kTable) / sizeof(kTable[0]));
    return kTable;
}
}  // namespace spx
