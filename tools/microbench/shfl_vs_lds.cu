// Do warp shuffles and shared-memory accesses share one data pipe on B200?  (The question behind "replace the 16x16
// shared-memory transposes between the radix-16 passes by shuffles".)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench/shfl_vs_lds tools/microbench/shfl_vs_lds.cu
// One CTA of 512 threads per SM.  Warps 0-7 form group S, warps 8-15 group L.
//   mode 0: group S runs K shuffles (32-bit, xor pattern), group L idles
//   mode 1: group L runs K/4 conflict-free LDS.128 + K/4 STS.128 ... scaled so that it moves the same number of 128-byte
//           wavefronts as K shuffles (one SHFL.32 = one 128-byte wavefront, one LDS.128 / STS.128 of a warp = four)
//   mode 2: both at once.  Separate pipes -> time(2) ~ max(time(0), time(1)); one pipe -> time(2) ~ time(0) + time(1).
// A 16x16 transpose of 16-byte elements among 16 lanes needs 4 butterfly steps x 8 elements x 4 words = 128 SHFL per thread
// (plus selects), against 16 STS.128 + 16 LDS.128 = 128 wavefronts per warp through shared memory: the same count.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int NT = 512;
__global__ void __launch_bounds__(NT) k(int mode, int K, unsigned long long *cyc, float *sink) {
    extern __shared__ float4 sm[];
    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 8192; i += NT) sm[i] = make_float4(i, 1, 2, 3);
    __syncthreads();
    float acc = tid;
    float a0 = tid, a1 = tid + 1, a2 = tid + 2, a3 = tid + 3, a4 = tid + 4, a5 = tid + 5, a6 = tid + 6, a7 = tid + 7;
    float4 v = make_float4(tid, 0, 0, 0);
    const long long t0 = clock64();
    if (w < 8) {
        if (mode == 0 || mode == 2) {
            // eight independent chains: throughput, not latency
#pragma unroll 4
            for (int i = 0; i < K / 8; ++i) {
                const int m = (i & 15) + 1;
                a0 = __shfl_xor_sync(0xffffffffu, a0, m); a1 = __shfl_xor_sync(0xffffffffu, a1, m);
                a2 = __shfl_xor_sync(0xffffffffu, a2, m); a3 = __shfl_xor_sync(0xffffffffu, a3, m);
                a4 = __shfl_xor_sync(0xffffffffu, a4, m); a5 = __shfl_xor_sync(0xffffffffu, a5, m);
                a6 = __shfl_xor_sync(0xffffffffu, a6, m); a7 = __shfl_xor_sync(0xffffffffu, a7, m);
            }
            acc = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
        }
    } else {
        if (mode == 1 || mode == 2) {
            float4 *p = sm + (w - 8) * 1024 + lane;     // each warp its own 16 KB: 8 KB read, 8 KB written; lanes 16 bytes apart
            float4 *q = p + 512;
#pragma unroll 8
            for (int i = 0; i < K / 8; ++i) {           // K/8 loads + K/8 stores = K/4 128-bit accesses = K wavefronts
                float4 a = p[(i & 15) * 32];
                if (i & 1) { v.x += a.x; } else { v.y += a.y; }
                q[(i & 15) * 32] = make_float4(i, v.z, 0.f, 1.f);
            }
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (tid == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);
    if (acc == 1.2345f || v.x == 1.2345f) sink[0] = acc + v.y;
    // per-warp durations: the slowest warp of either group
    __shared__ unsigned long long mx;
    if (tid == 0) mx = 0;
    __syncthreads();
    if (lane == 0) atomicMax(&mx, (unsigned long long)(t1 - t0));
    __syncthreads();
    if (tid == 0) cyc[blockIdx.x] = mx;
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    unsigned long long *cyc;
    float *sink;
    CK(cudaMalloc(&cyc, sms * sizeof(unsigned long long)));
    CK(cudaMalloc(&sink, 4));
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    const int K = 1 << 14;
    double t[3];
    for (int mode = 0; mode < 3; ++mode) {
        for (int it = 0; it < 2; ++it) k<<<sms, NT, 131072>>>(mode, K, cyc, sink);
        CK(cudaDeviceSynchronize());
        std::vector<unsigned long long> h(sms);
        CK(cudaMemcpy(h.data(), cyc, sms * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        double s = 0;
        for (auto c : h) s += (double)c;
        t[mode] = s / sms;
    }
    printf("K = %d wavefronts per warp, 8 warps per group, per SM (mean over %d SMs)\n", K, sms);
    printf("  shuffles only          : %9.0f cycles  (%.2f cycles per warp-SHFL per SM)\n", t[0], t[0] / (8.0 * K));
    printf("  LDS.128/STS.128 only   : %9.0f cycles  (%.2f cycles per 128-byte wavefront per SM)\n", t[1], t[1] / (8.0 * K));
    printf("  both groups at once    : %9.0f cycles  (sum %.0f, max %.0f)\n", t[2], t[0] + t[1], t[0] > t[1] ? t[0] : t[1]);
    printf("  => %s\n", t[2] > 0.8 * (t[0] + t[1]) ? "shuffles and shared-memory accesses take turns on ONE data pipe"
                                                  : "shuffles overlap shared-memory accesses (separate throughput)");
    return 0;
}
