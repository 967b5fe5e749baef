// A 16x16 transpose among the 16 threads of a half-warp WITHOUT shared memory: write tensor memory with one tcgen05 shape, read
// it back with another.  (The question behind "take the two warp-local exchanges of the 4096 kernel off the LSU pipe".)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/tmem_xchg tools/microbench/tmem_xchg.cu
// Thread (lane s, register column c) writes with .32x32b: lane s, column c.  A .16x256b load at lane half hh hands thread t the
// columns 8 i + 2 (t % 4) + {0, 1} of lanes t / 4 + 8 v + 16 hh.  One such step therefore moves the top two lane bits of the
// source into the register index, two register-index bits into the low two lane bits, and shifts the other lane bits up by two;
// two steps transpose (u3 u2 u1 u0 | m3 m2 m1 m0) for threads laid out as lane = 2 u + c -> lane = 16 c + m.  The store-side twin
// (.16x256b store, .32x32b load) is the exact inverse.  Each step is done in four 16-column quarters, so a warp needs only 16
// columns of tensor memory (2 KB).
// Checks the data movement, then times: TMEM exchange alone, shared-memory exchange alone, both alternating in every warp.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void st32x16(uint32_t a, const uint32_t (&r)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(a),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
                 "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void ld32x16(uint32_t a, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
                   "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(a) : "memory");
}
__device__ __forceinline__ void ld16x256x2(uint32_t a, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(a) : "memory");
}
__device__ __forceinline__ void st16x256x2(uint32_t a, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(a), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// x[e][f]: 16 elements of 4 words.  HI = true exchanges element bits 3:2 (quarter = bits 1:0), else bits 1:0 (quarter = 3:2).
template <bool HI>
__device__ __forceinline__ int eidx(int a, int h) { return HI ? 4 * a + h : 4 * h + a; }

// forward step: .32x32b store, .16x256b loads
template <bool HI>
__device__ __forceinline__ void step_fwd(uint32_t (&x)[16][4], uint32_t tx) {
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        uint32_t s[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) s[j] = x[eidx<HI>((j >> 1) & 3, h)][(j & 1) + 2 * (j >> 3)];   // column j: f0 = j&1, a = (j>>1)&3, f1 = j>>3
        st32x16(tx, s);
        wait_st();
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            uint32_t r[8];
            ld16x256x2(tx + ((uint32_t)(16 * hh) << 16), r);
            // r[4 i + 2 v + f0]: rep i = f1, v = lane + 8, f0 -> the new element's exchanged bits are (hh, v)
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int v = 0; v < 2; ++v)
#pragma unroll
                    for (int f0 = 0; f0 < 2; ++f0) x[eidx<HI>(2 * hh + v, h)][f0 + 2 * i] = r[4 * i + 2 * v + f0];
        }
        wait_ld();
    }
}
// inverse step: .16x256b stores, .32x32b load
template <bool HI>
__device__ __forceinline__ void step_inv(uint32_t (&x)[16][4], uint32_t tx) {
#pragma unroll
    for (int h = 0; h < 4; ++h) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            uint32_t r[8];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int v = 0; v < 2; ++v)
#pragma unroll
                    for (int f0 = 0; f0 < 2; ++f0) r[4 * i + 2 * v + f0] = x[eidx<HI>(2 * hh + v, h)][f0 + 2 * i];
            st16x256x2(tx + ((uint32_t)(16 * hh) << 16), r);
        }
        wait_st();
        uint32_t s[16];
        ld32x16(tx, s);
        wait_ld();
#pragma unroll
        for (int j = 0; j < 16; ++j) x[eidx<HI>((j >> 1) & 3, h)][(j & 1) + 2 * (j >> 3)] = s[j];
    }
}

constexpr int NT = 512;
// mode 0: check;  1: TMEM exchanges only;  2: shared-memory exchanges only;  3: both, alternating
template <int mode>
__global__ void __launch_bounds__(NT) k(int reps, uint32_t *err, unsigned long long *cyc, uint32_t *dump) {
    extern __shared__ uint4 sm[];
    __shared__ uint32_t tbase;
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    if (w == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tbase)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tx = tbase + ((uint32_t)(32 * (w & 3)) << 16) + (uint32_t)(16 * (w >> 2));   // this warp's 16 columns
    uint32_t x[16][4];
#pragma unroll
    for (int e = 0; e < 16; ++e)
#pragma unroll
        for (int f = 0; f < 4; ++f) x[e][f] = (uint32_t)((w << 16) | (lane << 8) | (e << 2) | f);
    if constexpr (mode == 0) {
        uint32_t bad = 0;
        step_fwd<true>(x, tx);
        if (blockIdx.x == 0 && w == 0 && dump) {
#pragma unroll
            for (int e = 0; e < 16; ++e)
#pragma unroll
                for (int f = 0; f < 4; ++f) dump[(lane * 16 + e) * 4 + f] = x[e][f];
        }
        step_fwd<false>(x, tx);
        // thread lane' = 16 c + m holds, in register u, what lane 2 u + c held in register m
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int f = 0; f < 4; ++f) {
                const int c = lane >> 4, m = lane & 15;
                const uint32_t want = (uint32_t)((w << 16) | ((2 * u + c) << 8) | (m << 2) | f);
                if (x[u][f] != want) ++bad;
            }
        step_inv<false>(x, tx);
        step_inv<true>(x, tx);
#pragma unroll
        for (int e = 0; e < 16; ++e)
#pragma unroll
            for (int f = 0; f < 4; ++f)
                if (x[e][f] != (uint32_t)((w << 16) | (lane << 8) | (e << 2) | f)) bad += 1000;
        if (bad) atomicAdd(err, bad);
    } else {
        // shared-memory exchange: the kernel's own pattern, 16 STS.128 + 16 LDS.128 per thread, 16-byte pad per 16 entries
        uint4 *wb = sm + w * 560;   // two columns of 256 entries, 16-byte pad per 16 entries, 64-byte skew between the columns
        __syncthreads();
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            if constexpr (mode == 1 || mode == 3) {
                step_fwd<true>(x, tx);
                step_fwd<false>(x, tx);
                step_inv<false>(x, tx);
                step_inv<true>(x, tx);
            }
            if constexpr (mode == 4) {       // forward-type steps only (.32x32b store, .16x256b loads)
                step_fwd<true>(x, tx);
                step_fwd<false>(x, tx);
                step_fwd<true>(x, tx);
                step_fwd<false>(x, tx);
            }
            if constexpr (mode == 5) {       // inverse-type steps only (.16x256b stores, .32x32b load)
                step_inv<false>(x, tx);
                step_inv<true>(x, tx);
                step_inv<false>(x, tx);
                step_inv<true>(x, tx);
            }
            if constexpr (mode == 2 || mode == 3) {
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    const int u = lane >> 1, c = lane & 1;
#pragma unroll
                    for (int m = 0; m < 16; ++m) {
                        const int e = 16 * m + u;
                        wb[c * 276 + e + (e >> 4)] = make_uint4(x[m][0], x[m][1], x[m][2], x[m][3]);
                    }
                    __syncwarp();
#pragma unroll
                    for (int m = 0; m < 16; ++m) {
                        const int e = 16 * u + m;
                        const uint4 v = wb[c * 276 + e + (e >> 4)];
                        x[m][0] = v.x; x[m][1] = v.y; x[m][2] = v.z; x[m][3] = v.w;
                    }
                    __syncwarp();
                }
            }
        }
        const long long t1 = clock64();
        __shared__ unsigned long long mx;
        if (tid == 0) mx = 0;
        __syncthreads();
        if (lane == 0) atomicMax(&mx, (unsigned long long)(t1 - t0));
        __syncthreads();
        if (tid == 0) cyc[blockIdx.x] = mx;
        uint32_t acc = 0;
#pragma unroll
        for (int e = 0; e < 16; ++e)
#pragma unroll
            for (int f = 0; f < 4; ++f) acc ^= x[e][f];
        if (acc == 0x12345678u) err[1] = acc;
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tbase), "r"(512));
}

int main() {
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    uint32_t *err, *dump;
    unsigned long long *cyc;
    CK(cudaMalloc(&err, 8));
    CK(cudaMemset(err, 0, 8));
    CK(cudaMalloc(&dump, 32 * 64 * 4));
    CK(cudaMalloc(&cyc, sms * 8));
    const size_t smem = 16 * 560 * sizeof(uint4);
    CK(cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<0><<<sms, NT, smem>>>(1, err, cyc, dump);
    CK(cudaDeviceSynchronize());
    uint32_t herr[2];
    CK(cudaMemcpy(herr, err, 8, cudaMemcpyDeviceToHost));
    printf("check: %u mismatches (0 = the two-step exchange is the 16x16 transpose and its twin the inverse)\n", herr[0]);
    if (herr[0]) {
        std::vector<uint32_t> h(32 * 64);
        CK(cudaMemcpy(h.data(), dump, h.size() * 4, cudaMemcpyDeviceToHost));
        for (int lane : {0, 1, 4, 5, 31}) {
            printf("after step 1, lane %d:", lane);
            for (int e = 0; e < 16; ++e) printf(" [l%u e%u f%u]", (h[(lane * 16 + e) * 4] >> 8) & 255, (h[(lane * 16 + e) * 4] >> 2) & 15, h[(lane * 16 + e) * 4] & 3);
            printf("\n");
        }
    }
    const int reps = 2000;
    for (int mode = 1; mode <= 5; ++mode) {
        for (int rr : {10, reps}) {
            if (mode == 1) k<1><<<sms, NT, smem>>>(rr, err, cyc, nullptr);
            if (mode == 2) k<2><<<sms, NT, smem>>>(rr, err, cyc, nullptr);
            if (mode == 3) k<3><<<sms, NT, smem>>>(rr, err, cyc, nullptr);
            if (mode == 4) k<4><<<sms, NT, smem>>>(rr, err, cyc, nullptr);
            if (mode == 5) k<5><<<sms, NT, smem>>>(rr, err, cyc, nullptr);
        }
        CK(cudaDeviceSynchronize());
        std::vector<unsigned long long> h(sms);
        CK(cudaMemcpy(h.data(), cyc, sms * 8, cudaMemcpyDeviceToHost));
        double s = 0;
        for (auto v : h) s += (double)v;
        s /= sms;
        // per rep: mode 1 = 2 exchanges (each two steps) of a 128 KB tile; mode 2 = 2 exchanges through shared memory
        printf("mode %d (%s): %.0f cycles per rep per SM = %.0f per 128 KB exchange\n", mode,
               mode == 1 ? "TMEM fwd+inv" : mode == 2 ? "smem x2" : mode == 3 ? "TMEM fwd+inv and smem x2 alternating" :
               mode == 4 ? "TMEM forward-type steps x4" : "TMEM inverse-type steps x4", s / reps, s / reps / (mode == 3 ? 4 : 2));
    }
    return 0;
}
