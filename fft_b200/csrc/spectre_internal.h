// Internal: helpers shared by the translation units of libspectre_mix.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

namespace spx {
// Record a message for spectre_mix_last_error() on this host thread and return `code` (defined in spectre_mix_api.cu).
int fail(int code, const char *fmt, ...);
// Same for a CUDA runtime error: clears the runtime's error state, returns SPECTRE_MIX_ERR_CUDA + e.
int cuda_fail(cudaError_t e, const char *what);
// Per-(device, F_half, Bk) interpolation table of the gate generator (spectre_gate.cu): built on first use (one small kernel
// + a stream synchronisation under a mutex -- like the twiddle tables, make one warm-up call before capturing a graph), read
// lock-free afterwards.  Returns 0 and the two device pointers, or an error code.
int gate_interp_table(int F_half, int Bk, const float4 **icoef, const ushort4 **itap);
}  // namespace spx
