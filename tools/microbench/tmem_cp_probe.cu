// tcgen05.cp (shared memory -> tensor memory, asynchronous, issued by one thread) probe for B200:
// which shared-memory entry lands in which TMEM lane / column for the 128x128b shape with a no-swizzle descriptor,
// and how long 128 KB take.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_cp_probe tmem_cp_probe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}
__global__ void __launch_bounds__(128) k(uint32_t *out, unsigned long long *ns, uint32_t lbo, uint32_t sbo, int reps) {
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ uint32_t tbase;
    __shared__ __align__(8) unsigned long long mbar;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    uint4 *buf = reinterpret_cast<uint4 *>(smem);           // 8192 entries of 16 B = 128 KB
    for (int e = tid; e < 8192; e += 128) buf[e] = make_uint4(4 * e, 4 * e + 1, 4 * e + 2, 4 * e + 3);
    if (w == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s32(&tbase)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes of buf -> visible to the async proxy
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = tbase;
    uint32_t phase = 0;
    unsigned long long t0 = 0, t1 = 0;
    // ---- semantics: first 8 KB (512 entries) -> columns 0..15
    if (tid == 0) {
        for (int g = 0; g < 4; ++g) {
            const uint64_t d = make_desc(s32(buf) + g * 2048, lbo, sbo);
            asm volatile("tcgen05.cp.cta_group::1.128x128b [%0], %1;" ::"r"(base + 4 * g), "l"(d) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&mbar)) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(s32(&mbar)), "r"(phase) : "memory");
    phase ^= 1;
    asm volatile("tcgen05.fence::after_thread_sync;");
    {
        const uint32_t quad = (uint32_t)(32 * w) << 16;
        for (int c = 0; c < 16; c += 4) {
            uint32_t r0, r1, r2, r3;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(base + quad + c) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;");
            uint32_t *o = out + (size_t)(32 * w + lane) * 16 + c;
            o[0] = r0; o[1] = r1; o[2] = r2; o[3] = r3;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    // ---- timing: 128 KB = 64 copies of 2 KB into 256 columns, reps times
    if (tid == 0) {
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
        for (int r = 0; r < reps; ++r) {
            for (int g = 0; g < 64; ++g) {
                const uint64_t d = make_desc(s32(buf) + g * 2048, lbo, sbo);
                asm volatile("tcgen05.cp.cta_group::1.128x128b [%0], %1;" ::"r"(base + 4 * g), "l"(d) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&mbar)) : "memory");
            asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(s32(&mbar)), "r"(phase) : "memory");
            phase ^= 1;
        }
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
        ns[blockIdx.x] = t1 - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512));
}
int main() {
    uint32_t *out; unsigned long long *ns;
    CK(cudaMalloc(&out, 128 * 16 * 4)); CK(cudaMalloc(&ns, 148 * 8));
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072 + 1024));
    const uint32_t combos[][2] = {{128, 128}, {2048, 128}};   // (an SBO other than 128 walks out of the 128 KB window)
    for (auto &cb : combos) {
        CK(cudaMemset(out, 0xff, 128 * 16 * 4));
        k<<<1, 128, 131072>>>(out, ns, cb[0], cb[1], 1);
        CK(cudaDeviceSynchronize());
        std::vector<uint32_t> h(128 * 16);
        CK(cudaMemcpy(h.data(), out, h.size() * 4, cudaMemcpyDeviceToHost));
        int bad = 0;
        for (int l = 0; l < 128; ++l) for (int c = 0; c < 16; ++c) bad += h[l * 16 + c] != (uint32_t)(((c / 4) * 128 + l) * 4 + c % 4);
        printf("lbo %u sbo %u: mismatches vs (entry e -> lane e%%128, columns 4*(e/128)..+3) = %d | lane0: %u %u %u %u %u | lane1: %u %u | lane8: %u %u | lane 32: %u\n",
               cb[0], cb[1], bad, h[0], h[1], h[2], h[3], h[4], h[16], h[17], h[128], h[129], h[512]);
    }
    const int reps = 200;
    k<<<148, 128, 131072>>>(out, ns, 128, 128, reps);
    CK(cudaDeviceSynchronize());
    std::vector<unsigned long long> t(148);
    CK(cudaMemcpy(t.data(), ns, 148 * 8, cudaMemcpyDeviceToHost));
    double s = 0; for (auto x : t) s += (double)x;
    printf("tcgen05.cp of 128 KB per SM (64 x 2 KB, one thread, commit + wait each tile): %.0f ns per tile\n", s / 148 / reps);
    return 0;
}
