// Internal: helpers shared by the translation units of libspectre_mix.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

namespace spx {
// Record a message for spectre_mix_last_error() on this host thread and return `code` (defined in spectre_mix_api.cu).
int fail(int code, const char *fmt, ...);
// Same for a CUDA runtime error: clears the runtime's error state, returns SPECTRE_MIX_ERR_CUDA + e.
int cuda_fail(cudaError_t e, const char *what);
}  // namespace spx
