// Internal: pre/post streaming passes of the long-context two-pass path (spectre_long.cu).
#pragma once
#include <cuda_runtime.h>

namespace spx {
// pre = true:  user V ([B][rows][C], strides u_sb/u_sn in elements, fp32 or bf16) -> scratch fp32 [B][R][sub][C]
// pre = false: scratch -> user out ([B][rows][C])
// twid = device table of W_{R sub}^{u q}, [q - 1][u], q = 1 .. R-1, u < sub (built by the API layer, cached per device and R)
cudaError_t long_pass(bool pre, int R, int dtype_bf16, const void *src, void *dst, long long u_sb, long long u_sn, int B, int rows,
                      int C, int sub, int sms, const float2 *twid, cudaStream_t st);
}  // namespace spx
