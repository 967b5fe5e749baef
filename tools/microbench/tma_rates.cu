// TMA row-rate microbenchmark (B200): how fast can one SM pull / push [rows x W bytes] boxes of a strided tensor?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_rates tma_rates.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const void *tm, int c, int r, int b, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(tm), "r"(c), "r"(r), "r"(b), "r"(bar) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const void *tm, uint32_t src, int c, int r, int b) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(tm), "r"(c), "r"(r), "r"(b), "r"(src) : "memory");
}

// mode 0: loads only; 1: stores only; 2: load tile i+1 while storing tile i (issued back to back by one thread);
// mode 3: like 2 but loads and stores issued by two different threads (warps)
template <int MODE>
__global__ void k(const __grid_constant__ CUtensorMap tin, const __grid_constant__ CUtensorMap tout, int rows, int wbytes,
                  int tiles_per_row, int num_tiles, int boxr, unsigned long long *out_ns) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[2];
    const uint32_t bar = smem_u32(&bars[0]);
    const int tile_bytes = rows * wbytes;
    unsigned char *buf0 = smem, *buf1 = smem;   // contents are irrelevant here: stores read the landing buffer
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    unsigned long long t0 = 0;
    if (threadIdx.x == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    uint32_t par = 0;
    const int cw = wbytes / 4;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int b = tile / tiles_per_row, c = (tile % tiles_per_row) * cw;
        if (MODE == 0 || MODE == 2 || MODE == 3) {
            if (threadIdx.x == 0) {
                mbar_expect_tx(bar, tile_bytes);
                for (int r0 = 0; r0 < rows; r0 += boxr) tma_load_3d(smem_u32(buf0) + r0 * wbytes, &tin, c, r0, b, bar);
            }
        }
        if (MODE == 1 || MODE == 2) {
            if (threadIdx.x == 0) {
                for (int r0 = 0; r0 < rows; r0 += boxr) tma_store_3d(&tout, smem_u32(buf1) + r0 * wbytes, c, r0, b);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if (MODE == 3) {
            if (threadIdx.x == 32) {
                for (int r0 = 0; r0 < rows; r0 += boxr) tma_store_3d(&tout, smem_u32(buf1) + r0 * wbytes, c, r0, b);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
        }
        if (MODE == 1 || MODE == 2) {
            if (threadIdx.x == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        if (MODE == 0 || MODE == 2 || MODE == 3) {
            mbar_wait(bar, par);
            par ^= 1;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0 || threadIdx.x == 32) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
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
    int sms = 148;
    if (getenv("TMA_SMS")) sms = atoi(getenv("TMA_SMS"));
    for (int wbytes : {32, 64, 128}) {
        for (int boxr : {256, 64}) {
            const int rows = 128 * 1024 / wbytes / ((wbytes == 32) ? 1 : 1);  // 128 KB tiles
            const int trows = rows > N ? N : rows;
            if (trows < boxr) continue;
            CUtensorMap tin, tout;
            cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)N, (cuuint64_t)B};
            cuuint64_t strides[2] = {(cuuint64_t)C * 4, (cuuint64_t)N * C * 4};
            cuuint32_t box[3] = {(cuuint32_t)(wbytes / 4), (cuuint32_t)boxr, 1};
            cuuint32_t es[3] = {1, 1, 1};
            if (enc(&tin, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, in, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) ||
                enc(&tout, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, out, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)) {
                printf("encode failed\n");
                return 1;
            }
            const int tiles_per_row = C * 4 / wbytes;
            const int tile_bytes = trows * wbytes;
            // tiles cover rows [0, trows) only of each batch row: num tiles = B * tiles_per_row
            const int num_tiles = B * tiles_per_row;
            const size_t smem = (size_t)tile_bytes;
            for (int mode = 0; mode < 4; ++mode) {
                auto fn = mode == 0 ? k<0> : mode == 1 ? k<1> : mode == 2 ? k<2> : k<3>;
                CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                for (int rep = 0; rep < 2; ++rep) {
                    fn<<<sms, 64, smem>>>(tin, tout, trows, wbytes, tiles_per_row, num_tiles, boxr, ns);
                    CK(cudaDeviceSynchronize());
                }
                std::vector<unsigned long long> h(sms);
                CK(cudaMemcpy(h.data(), ns, sms * 8, cudaMemcpyDeviceToHost));
                double avg = 0;
                for (auto v : h) avg += v;
                avg /= sms;
                const double tiles_per_cta = (double)num_tiles / sms;
                const double us_per_tile = avg / 1e3 / tiles_per_cta;
                const double bytes = (double)num_tiles * tile_bytes * ((mode >= 2) ? 2 : 1);
                printf("row=%3dB box=%3d rows tile=%3dKB mode=%d (%s): %.2f us/tile/SM, %.0f GB/s aggregate, %.2f cycles/row@1.9GHz\n",
                       wbytes, boxr, tile_bytes / 1024, mode,
                       mode == 0 ? "load" : mode == 1 ? "store" : mode == 2 ? "load+store 1 thr" : "load+store 2 thr", us_per_tile,
                       bytes / (avg * 1e-9) / 1e9, us_per_tile * 1900.0 / trows / ((mode >= 2) ? 2 : 1));
            }
        }
    }
    return 0;
}
