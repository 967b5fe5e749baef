// TMEM as a staging memory: semantics + bandwidth probe (B200).
//  * warp w of a CTA reaches TMEM lanes 32*(w%4) .. +31; thread i of the warp = lane base+i; address = lane<<16 | column
//  * writer warp w writes pattern, reader warp w+4 (same quadrant) reads it back after a CTA barrier
//  * then times 128 KB of tcgen05.st and tcgen05.ld per CTA (16 warps x 32 lanes x 64 columns x 4 B)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)
__device__ __forceinline__ void tst4(uint32_t a, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
}
__device__ __forceinline__ void tld4(uint32_t a, uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a) : "memory");
}
__global__ void __launch_bounds__(512) k(uint32_t *err, unsigned long long *ns, int reps) {
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    if (w == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = tbase;
    const uint32_t quad = (uint32_t)(32 * (w & 3)) << 16;
    // ---- semantics: warps 0..3 write columns [0,64) of their quadrant, warps 4..7 read them back
    if (w < 4) {
        for (int c = 0; c < 64; c += 4) {
            const uint32_t v = (uint32_t)((32 * w + lane) * 1000 + c);
            tst4(base + quad + c, v, v + 1, v + 2, v + 3);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (w >= 4 && w < 8) {
        uint32_t bad = 0;
        for (int c = 0; c < 64; c += 4) {
            uint32_t r0, r1, r2, r3;
            tld4(base + quad + c, r0, r1, r2, r3);
            asm volatile("tcgen05.wait::ld.sync.aligned;");
            const uint32_t v = (uint32_t)((32 * (w & 3) + lane) * 1000 + c);
            bad += (r0 != v) + (r1 != v + 1) + (r2 != v + 2) + (r3 != v + 3);
        }
        if (bad) atomicAdd(err, bad);
    }
    __syncthreads();
    // ---- bandwidth: every warp stores / loads 16 x (x4 columns) = 64 columns = 8 KB per warp, 128 KB per CTA
    const uint32_t colbase = (uint32_t)(64 * (w >> 2));   // 4 warps per quadrant share the 512 columns... 4*64=256 used
    unsigned long long t0, t1, t2;
    __syncthreads();
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int c = 0; c < 64; c += 4) tst4(base + quad + colbase + c, r, r + 1, r + 2, r + 3);
        asm volatile("tcgen05.wait::st.sync.aligned;");
    }
    __syncthreads();
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    uint32_t acc = 0;
    for (int r = 0; r < reps; ++r) {
        uint32_t v[64];
#pragma unroll
        for (int c = 0; c < 64; c += 4) tld4(base + quad + colbase + c, v[c], v[c + 1], v[c + 2], v[c + 3]);
        asm volatile("tcgen05.wait::ld.sync.aligned;");
#pragma unroll
        for (int c = 0; c < 64; ++c) acc += v[c];
    }
    __syncthreads();
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t2));
    if (acc == 0x12345678) atomicAdd(err, 1);
    if (tid == 0) {
        ns[blockIdx.x * 2] = t1 - t0;
        ns[blockIdx.x * 2 + 1] = t2 - t1;
    }
    __syncthreads();
    if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(512));
}
int main() {
    uint32_t *err;
    unsigned long long *ns;
    CK(cudaMalloc(&err, 4));
    CK(cudaMemset(err, 0, 4));
    CK(cudaMalloc(&ns, 148 * 16));
    const int reps = 200;
    k<<<148, 512>>>(err, ns, reps);
    CK(cudaDeviceSynchronize());
    uint32_t h;
    CK(cudaMemcpy(&h, err, 4, cudaMemcpyDeviceToHost));
    std::vector<unsigned long long> t(296);
    CK(cudaMemcpy(t.data(), ns, 296 * 8, cudaMemcpyDeviceToHost));
    double st = 0, ld = 0;
    for (int i = 0; i < 148; ++i) { st += t[2 * i]; ld += t[2 * i + 1]; }
    st /= 148 * reps; ld /= 148 * reps;
    printf("mismatches %u; 128 KB per CTA: tcgen05.st %.0f ns (%.0f B/clk @1.9GHz), tcgen05.ld %.0f ns (%.0f B/clk)\n", h, st,
           131072.0 / (st * 1.9), ld, 131072.0 / (ld * 1.9));
    return 0;
}
