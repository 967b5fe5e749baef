// Kernel variants for n_fft group "16384": one table per file so that the instantiations compile in parallel (make -j).
// Registry order matters: the first entry of a (n_fft, dtype, mode) that fits is the default (spectre_mix_api.cu: choose()).
#include "spectre_mix_registry.h"
#include "../../include/spectre_mix.h"

namespace spx {
namespace {
const KernelEntry kTable[] = {
    SPX_ENTRY(4, 16, 16, 16, MODE_PAIR, 1, 512, 1, float, SPECTRE_MIX_F32),
    SPX_ENTRY(4, 16, 16, 16, MODE_PAIR, 1, 512, 1, __nv_bfloat16, SPECTRE_MIX_BF16),
    SPX_ENTRY(4, 16, 16, 16, MODE_REAL, 1, 512, 1, float, SPECTRE_MIX_F32),
    SPX_ENTRY(4, 16, 16, 16, MODE_REAL, 1, 512, 1, __nv_bfloat16, SPECTRE_MIX_BF16),
};
}  // namespace
const KernelEntry *table_16384(int *count) {
    *count = (int)(sizeof(kTable) / sizeof(kTable[0]));
    return kTable;
}
}  // namespace spx
