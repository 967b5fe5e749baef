// Kernel variants for n_fft group "8192": one table per file so that the instantiations compile in parallel (make -j).
// Registry order matters: the first entry of a (n_fft, dtype, mode) that fits is the default (spectre_mix_api.cu: choose()).
#include "spectre_mix_registry.h"
#include "../../include/spectre_mix.h"

namespace spx {
namespace {
const KernelEntry kTable[] = {
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
