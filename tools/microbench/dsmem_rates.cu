// Distributed-shared-memory exchange rates inside a thread-block cluster (B200), the question behind a one-HBM-pass
// long-context kernel: a 2-/4-CTA cluster that exchanges the radix-R stage of a 8192/16384-point transform has to move
// (R-1)/R of every 128 KB tile slab to the peer CTAs, twice per tile.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/dsmem_rates tools/microbench/dsmem_rates.cu
// Every CTA of every cluster (all SMs busy) moves `KB` KB per repetition to its peers:
//   mode 0  local STS.128 (baseline)            mode 1  st.shared::cluster.v4 to the peers, round robin
//   mode 2  ld.shared::cluster.v4 from peers    mode 3  cp.async.bulk shared::cta -> shared::cluster (TMA engine), 16 KB pieces
// Timed with clock64 per CTA between cluster barriers; prints cycles per repetition and bytes per clock per SM.
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, float4 v) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ld_cluster_v4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

constexpr int NT = 512;
constexpr int BUF = 96 * 1024;   // bytes exchanged per repetition and CTA (= 3/4 of a 128 KB slab)

template <int MODE>
__global__ void __launch_bounds__(NT) k(unsigned long long *cyc, float *sink, int reps, int csize) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float4 *buf = reinterpret_cast<float4 *>(smem);                 // [BUF] source / landing area
    float4 *src = reinterpret_cast<float4 *>(smem + BUF);           // [BUF/2 .. ] only for mode 3 (source of bulk copies)
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 2 * BUF);
    cg::cluster_group cl = cg::this_cluster();
    const uint32_t rank = cl.block_rank();
    const int tid = threadIdx.x;
    constexpr int N16 = BUF / 16;
    for (int i = tid; i < N16; i += NT) { buf[i] = make_float4(i, rank, 0, 0); if (MODE == 3) src[i] = make_float4(i, 1, 2, 3); }
    if (MODE == 3 && tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    cl.sync();
    float4 acc = make_float4(0, 0, 0, 0);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
        if (MODE == 0) {
#pragma unroll 4
            for (int i = tid; i < N16; i += NT) buf[i] = make_float4(r, i, 0, 0);
        } else if (MODE == 1) {
            // piece p of the buffer goes to peer (rank + 1 + p % (csize-1)) % csize; consecutive lanes, consecutive 16-byte entries
            const int per = N16 / (csize - 1);
#pragma unroll 4
            for (int i = tid; i < N16; i += NT) {
                const uint32_t peer = (rank + 1 + i / per) % csize;
                st_cluster_v4(mapa(smem_u32(buf + i), peer), make_float4(r, i, 0, 0));
            }
        } else if (MODE == 2) {
            const int per = N16 / (csize - 1);
#pragma unroll 4
            for (int i = tid; i < N16; i += NT) {
                const uint32_t peer = (rank + 1 + i / per) % csize;
                const float4 v = ld_cluster_v4(mapa(smem_u32(buf + i), peer));
                acc.x += v.x; acc.y += v.y;
            }
        } else {
            // TMA engine: BUF bytes as 16 KB bulk copies from local `src` to the peers' `buf`, completion on the PEER's mbarrier
            constexpr int PIECE = 16 * 1024, NP = BUF / PIECE;
            if (tid == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(BUF) : "memory");
            }
            cl.sync();   // every CTA's barrier is armed before any peer copies into it
            if (tid == 0) {
                for (int p = 0; p < NP; ++p) {
                    const uint32_t peer = (rank + 1 + p % (csize - 1)) % csize;
                    const uint32_t dst = mapa(smem_u32(smem + (size_t)p * PIECE), peer), rbar = mapa(smem_u32(bar), peer);
                    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                                 "r"(smem_u32(smem + BUF + (size_t)p * PIECE)), "r"(PIECE), "r"(rbar)
                                 : "memory");
                }
            }
            // wait until MY buffer has received BUF bytes from the peers
            asm volatile(
                "{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(
                    smem_u32(bar)),
                "r"(r & 1)
                : "memory");
        }
        if (MODE != 3) cl.sync();
    }
    const long long t1 = clock64();
    if (tid == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc.x == 123.456f) sink[0] = acc.y;
    cl.sync();
}

template <int MODE>
void run(int csize, int reps, const char *name) {
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = sms / csize * csize;
    const size_t smem = 2 * BUF + 64;
    CK(cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    unsigned long long *cyc;
    float *sink;
    CK(cudaMalloc(&cyc, grid * sizeof(unsigned long long)));
    CK(cudaMalloc(&sink, 4));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NT);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    for (int it = 0; it < 2; ++it) CK(cudaLaunchKernelEx(&cfg, k<MODE>, cyc, sink, reps, csize));
    CK(cudaDeviceSynchronize());
    std::vector<unsigned long long> h(grid);
    CK(cudaMemcpy(h.data(), cyc, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    double mx = 0, sum = 0;
    for (auto c : h) { sum += (double)c; if ((double)c > mx) mx = (double)c; }
    const double mean = sum / grid / reps;
    printf("%-34s cluster %d  grid %3d: %8.0f cycles per 96 KB (max CTA %8.0f) = %6.1f B/clk/SM\n", name, csize, grid, mean, mx / reps,
           BUF / mean);
    cudaFree(cyc); cudaFree(sink);
}

int main() {
    const int reps = 50;
    run<0>(2, reps, "local STS.128 (+cluster barrier)");
    for (int cs : {2, 4}) {
        run<1>(cs, reps, "st.shared::cluster.v4 to peers");
        run<2>(cs, reps, "ld.shared::cluster.v4 from peers");
        run<3>(cs, reps, "cp.async.bulk smem->peer smem 16KB");
    }
    return 0;
}
