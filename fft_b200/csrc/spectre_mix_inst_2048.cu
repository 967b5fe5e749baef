// Kernel variants for n_fft group "2048": one table per file so that the instantiations compile in parallel (make -j).
// Registry order matters: the first entry of a (n_fft, dtype, mode) that fits is the default (spectre_mix_api.cu: choose()).
#include "spectre_mix_registry.h"
#include "../../include/spectre_mix.h"

namespace spx {
namespace {
const KernelEntry kTable[] = {
    // wide-row TMEM-staged variant (default): 2048 = 16 x 16 x 8, 16-channel tiles = 64-byte rows, 128 KB per tile
    SPX_ENTRY(16, 16, 8, 1, MODE_QUAD, 4, 512, 1, float, SPECTRE_MIX_F32),
    SPX_ENTRY(16, 16, 8, 1, MODE_QUAD, 4, 512, 1, __nv_bfloat16, SPECTRE_MIX_BF16),
    SPX_ENTRY(8, 16, 16, 1, MODE_QUAD, 2, 256, 2, float, SPECTRE_MIX_F32),
    SPX_ENTRY(8, 16, 16, 1, MODE_QUAD, 2, 256, 2, __nv_bfloat16, SPECTRE_MIX_BF16),
    SPX_ENTRY(8, 16, 16, 1, MODE_QUAD, 4, 512, 1, float, SPECTRE_MIX_F32),             // 16-channel tile, one CTA per SM
    SPX_ENTRY(8, 16, 16, 1, MODE_QUAD, 4, 512, 1, __nv_bfloat16, SPECTRE_MIX_BF16),
    SPX_ENTRY(8, 16, 16, 1, MODE_QUAD, 1, 128, 4, float, SPECTRE_MIX_F32),
    SPX_ENTRY(8, 16, 16, 1, MODE_QUAD, 1, 128, 4, __nv_bfloat16, SPECTRE_MIX_BF16),
    SPX_ENTRY(8, 16, 16, 1, MODE_PAIR, 4, 512, 1, float, SPECTRE_MIX_F32),
    SPX_ENTRY(8, 16, 16, 1, MODE_PAIR, 4, 512, 1, __nv_bfloat16, SPECTRE_MIX_BF16),
    SPX_ENTRY(8, 16, 16, 1, MODE_REAL, 4, 512, 1, float, SPECTRE_MIX_F32),
    SPX_ENTRY(8, 16, 16, 1, MODE_REAL, 4, 512, 1, __nv_bfloat16, SPECTRE_MIX_BF16),
    SPX_ENTRY(8, 16, 16, 1, MODE_PAIR, 1, 128, 2, float, SPECTRE_MIX_F32),   // narrow fallback: one gate group per tile
    SPX_ENTRY(8, 16, 16, 1, MODE_PAIR, 1, 128, 2, __nv_bfloat16, SPECTRE_MIX_BF16),   // narrow fallback: one gate group per tile
    SPX_ENTRY(8, 16, 16, 1, MODE_REAL, 1, 128, 2, float, SPECTRE_MIX_F32),   // narrow fallback: one gate group per tile
    SPX_ENTRY(8, 16, 16, 1, MODE_REAL, 1, 128, 2, __nv_bfloat16, SPECTRE_MIX_BF16),   // narrow fallback: one gate group per tile
};
}  // namespace
const KernelEntry *table_2048(int *count) {
    *count = (int)(sizeof(kTable) / sizeof(kTable[0]));
    return kTable;
}
}  // namespace spx
