// Can the LSU path (cp.async 16B / LDG) bring a tile in while the TMA unit pushes the previous tile out?
// 32-byte rows (8 fp32 channels of a [B][4096][768] tensor), 128 KB tiles, per-SM rates with few or all SMs active.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tma_store_3d(const void *tm, uint32_t src, int c, int r, int b) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(tm), "r"(c), "r"(r), "r"(b), "r"(src) : "memory");
}
// mode 0: cp.async loads only; 1: TMA stores only; 2: cp.async loads of tile i+1 overlapped with TMA stores of tile i;
// mode 3: LDG->STS loads only; 4: LDG->STS loads overlapped with TMA stores
template <int MODE>
__global__ void __launch_bounds__(512) k(const float *in, const __grid_constant__ CUtensorMap tout, int N, int C, int tiles_per_row,
                                         int num_tiles, unsigned long long *out_ns) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float4 *buf = reinterpret_cast<float4 *>(smem);
    const int tid = threadIdx.x;
    unsigned long long t0 = 0;
    __syncthreads();
    if (tid == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int b = tile / tiles_per_row, c = (tile % tiles_per_row) * 8;
        if (MODE == 1 || MODE == 2 || MODE == 4) {
            if (tid == 0) {
                for (int r0 = 0; r0 < N; r0 += 256) tma_store_3d(&tout, smem_u32(buf) + r0 * 32, c, r0, b);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if (MODE == 0 || MODE == 2) {
            const float *src = in + ((size_t)b * N) * C + c;
            // thread -> (row = tid/2 + 256 m, half = tid&1): 16 cp.async of 16 B each
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const int row = (tid >> 1) + 256 * m;
                const float *g = src + (size_t)row * C + (tid & 1) * 4;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(buf + row * 2 + (tid & 1))), "l"(g) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        if (MODE == 3 || MODE == 4) {
            const float *src = in + ((size_t)b * N) * C + c;
            float4 v[16];
#pragma unroll
            for (int m = 0; m < 16; ++m) {
                const int row = (tid >> 1) + 256 * m;
                v[m] = __ldcs(reinterpret_cast<const float4 *>(src + (size_t)row * C + (tid & 1) * 4));
            }
#pragma unroll
            for (int m = 0; m < 16; ++m) buf[((tid >> 1) + 256 * m) * 2 + (tid & 1) + 8192] = v[m];   // second half of smem
        }
        if (MODE == 1 || MODE == 2 || MODE == 4) {
            if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncthreads();
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
        out_ns[blockIdx.x] = t1 - t0;
    }
}
typedef CUresult (*EncFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                          const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
    const int B = 64, N = 4096, C = 768;
    float *in, *out;
    CK(cudaMalloc(&in, (size_t)B * N * C * 4));
    CK(cudaMalloc(&out, (size_t)B * N * C * 4));
    CK(cudaMemset(in, 0, (size_t)B * N * C * 4));
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncFn enc = (EncFn)p;
    unsigned long long *ns;
    CK(cudaMalloc(&ns, 148 * 8));
    CUtensorMap tout;
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)N, (cuuint64_t)B};
    cuuint64_t strides[2] = {(cuuint64_t)C * 4, (cuuint64_t)N * C * 4};
    cuuint32_t box[3] = {8, 256, 1};
    cuuint32_t es[3] = {1, 1, 1};
    if (enc(&tout, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) return 1;
    const int tiles_per_row = C / 8, num_tiles = B * tiles_per_row;
    const size_t smem = 2 * 128 * 1024 > 220 * 1024 ? 220 * 1024 : 2 * 128 * 1024;
    for (int sms : {8, 148}) {
        for (int mode = 0; mode < 5; ++mode) {
            auto fn = mode == 0 ? k<0> : mode == 1 ? k<1> : mode == 2 ? k<2> : mode == 3 ? k<3> : k<4>;
            const size_t sm = (mode >= 3) ? 220 * 1024 : 128 * 1024;
            CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            for (int rep = 0; rep < 2; ++rep) {
                fn<<<sms, 512, sm>>>(in, tout, N, C, tiles_per_row, num_tiles, ns);
                CK(cudaDeviceSynchronize());
            }
            std::vector<unsigned long long> h(sms);
            CK(cudaMemcpy(h.data(), ns, sms * 8, cudaMemcpyDeviceToHost));
            double avg = 0;
            for (auto v : h) avg += v;
            avg /= sms;
            const double us_per_tile = avg / 1e3 / ((double)num_tiles / sms);
            const char *names[] = {"cp.async 16B loads", "TMA stores", "cp.async loads || TMA stores", "LDG.128 loads", "LDG loads || TMA stores"};
            printf("SMs=%3d %-30s %.2f us per 128 KB tile per SM\n", sms, names[mode], us_per_tile);
        }
    }
    (void)smem;
    return 0;
}
