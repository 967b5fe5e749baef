// How fast can ONE shared-memory FFT pass of the 4096 x 8-channel tile run in isolation (no HBM traffic, no helper warps)?
// Runs the kernel's own pass code (fwd_inner_pass stage 1 = 16 LDS.128 + 15 twiddle LDS.64 + radix-16 packed butterfly
// + 16 STS.128 per thread, 512 threads) back to back and compares with its two halves:
//   mode 0: full pass   mode 1: shared-memory traffic only (no butterfly)   mode 2: butterfly only (registers)
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../fft_b200/csrc/spectre_mix_kernel.cuh"
using namespace spx;
using PL = Plan<16, 16, 16, 1>;
constexpr int NT = 512, NCOL = 2;
using SM = Smem<PL, MODE_QUAD, NCOL>;

// one thread = the same butterfly of BOTH element columns: loads of both first, then butterfly/store of column 0 while
// column 1's loads land, butterfly/store of column 1 while column 0's stores drain; twiddles fetched once for both
__device__ __forceinline__ float4 lds_v(const float4 *p) {
    float4 v;
    asm volatile("ld.volatile.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
    return v;
}
__device__ __forceinline__ void sts_v(float4 *p, float4 v) {
    asm volatile("st.volatile.shared.v4.f32 [%4], {%0,%1,%2,%3};" ::"f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "r"(smem_u32(p)));
}
// a one-warp named barrier: completes at once, but ptxas schedules nothing across it
__device__ __forceinline__ void region_fence(int tid) { asm volatile("bar.sync %0, 32;" ::"r"(4 + (tid >> 5)) : "memory"); }
template <int SYNC, int EARLY>
__global__ void __launch_bounds__(256, 1) kdual(int reps, const float2 *twg, unsigned long long *out, float *sink) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float4 *buf = reinterpret_cast<float4 *>(smem_raw);
    float2 *tw = reinterpret_cast<float2 *>(smem_raw + SM::data_bytes);
    const int tid = threadIdx.x;
    for (int i = tid; i < (int)(SM::data_bytes / 16); i += 256) buf[i] = make_float4(0.001f * i, 1.f, 0.5f, 0.25f);
    for (int i = tid; i < PL::TWN; i += 256) tw[i] = make_float2(0.8f, 0.6f);
    __syncthreads();
    unsigned long long t0 = clock64();
    constexpr int R = 16, L = 16, CS = SM::CS;
    using E = Elem<MODE_QUAD>;
    for (int r = 0; r < reps; ++r) {
        const int bf = tid;
        const int Q = bf / L, u = bf - Q * L, e0 = Q * (R * L) + u;
        float4 *cb = buf + e0 + (e0 >> 4);
        Cx<float2> x[R], y[R];
        auto ldx = [&](int m) { x[m] = E::unpack(lds_v(cb + m * L + ((m * L) >> 4))); };
        auto ldy = [&](int m) { y[m] = E::unpack(lds_v(cb + CS + m * L + ((m * L) >> 4))); };
        if (EARLY == 3) {
            // volatile accesses keep program order: column 1's loads are interleaved with column 0's (so they are issued before
            // any butterfly), column 0's stores come before column 1's last loads (so they are issued before its butterfly)
#pragma unroll
            for (int m = 0; m < 12; ++m) { ldx(m); ldy(m); }
#pragma unroll
            for (int m = 12; m < 16; ++m) ldx(m);
        } else {
            constexpr int XA = EARLY == 1 ? 12 : 16;
#pragma unroll
            for (int m = 0; m < XA; ++m) ldx(m);
#pragma unroll
            for (int m = 0; m < R; ++m) ldy(m);
#pragma unroll
            for (int m = XA; m < R; ++m) ldx(m);
        }
        float2 wq[R - 1];
#pragma unroll
        for (int q = 1; q < R; ++q) wq[q - 1] = tw[PL::TWOFF(1) + (q - 1) * L + u];
        if (EARLY == 2) region_fence(tid);
        Dft<R, float2>::run(x);
#pragma unroll
        for (int q = 1; q < R; ++q) x[q] = cmul(x[q], wq[q - 1].x, wq[q - 1].y);
#pragma unroll
        for (int q = 0; q < R; ++q) sts_v(cb + q * L + ((q * L) >> 4), E::pack(x[q]));
        if (EARLY == 2) region_fence(tid);
        if (EARLY == 3) {
#pragma unroll
            for (int m = 12; m < 16; ++m) ldy(m);
        }
        Dft<R, float2>::run(y);
#pragma unroll
        for (int q = 1; q < R; ++q) y[q] = cmul(y[q], wq[q - 1].x, wq[q - 1].y);
#pragma unroll
        for (int q = 0; q < R; ++q) sts_v(cb + CS + q * L + ((q * L) >> 4), E::pack(y[q]));
        if (SYNC) __syncthreads(); else __syncwarp();
    }
    __syncthreads();
    unsigned long long t1 = clock64();
    if (tid == 0) out[blockIdx.x] = t1 - t0;
    if (buf[tid].x == 12345.678f) sink[1] = buf[tid].x;
}

template <int MODE_>
__global__ void __launch_bounds__(NT, 1) k(int reps, const float2 *twg, unsigned long long *out, float *sink) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float4 *buf = reinterpret_cast<float4 *>(smem_raw);
    float2 *tw = reinterpret_cast<float2 *>(smem_raw + SM::data_bytes);
    const int tid = threadIdx.x;
    for (int i = tid; i < (int)(SM::data_bytes / 16); i += NT) buf[i] = make_float4(0.001f * i, 1.f, 0.5f, 0.25f);
    for (int i = tid; i < PL::TWN; i += NT) tw[i] = make_float2(0.8f, 0.6f);
    __syncthreads();
    unsigned long long t0 = clock64();
    if (MODE_ == 0) {
        for (int r = 0; r < reps; ++r) { fwd_inner_pass<PL, MODE_QUAD, NCOL, NT, 1>(buf, tw, twg, tid); __syncwarp(); }
    } else if (MODE_ == 3) {          // CTA barrier after every pass: all 16 warps start each pass in lock step
        for (int r = 0; r < reps; ++r) { fwd_inner_pass<PL, MODE_QUAD, NCOL, NT, 1>(buf, tw, twg, tid); __syncthreads(); }
    } else if (MODE_ == 4) {          // two 8-warp groups with their own barriers, the second one started half a pass later
        if (tid >= 256) { const unsigned long long w0 = clock64(); while (clock64() - w0 < 1300) {} }
        for (int r = 0; r < reps; ++r) {
            fwd_inner_pass<PL, MODE_QUAD, NCOL, NT, 1>(buf, tw, twg, tid);
            if (tid < 256) asm volatile("bar.sync 1, 256;" ::: "memory"); else asm volatile("bar.sync 2, 256;" ::: "memory");
        }
    } else if (MODE_ == 5) {          // two groups with their own barriers, no initial offset
        for (int r = 0; r < reps; ++r) {
            fwd_inner_pass<PL, MODE_QUAD, NCOL, NT, 1>(buf, tw, twg, tid);
            if (tid < 256) asm volatile("bar.sync 1, 256;" ::: "memory"); else asm volatile("bar.sync 2, 256;" ::: "memory");
        }
    } else if (MODE_ == 1) {
        constexpr int R = 16, L = 16, NBF = PL::N / R, CS = SM::CS;
        for (int r = 0; r < reps; ++r) {
            const int col = tid / NBF, bf = tid - col * NBF;
            const int Q = bf / L, u = bf - Q * L, e0 = Q * (R * L) + u;
            float4 *cb = buf + col * CS + e0 + (e0 >> 4);
            float4 x[16];
#pragma unroll
            for (int m = 0; m < R; ++m) x[m] = cb[m * L + ((m * L) >> 4)];
            float2 acc = make_float2(0.f, 0.f);
#pragma unroll
            for (int q = 1; q < R; ++q) { const float2 w = tw[PL::TWOFF(1) + (q - 1) * L + u]; acc.x += w.x; acc.y += w.y; }
#pragma unroll
            for (int q = 0; q < R; ++q) { x[q].x += acc.x; cb[q * L + ((q * L) >> 4)] = x[q]; }
            __syncwarp();
        }
    } else {
        Cx<float2> x[16];
#pragma unroll
        for (int m = 0; m < 16; ++m) x[m] = {make_float2(tid * 0.01f + m, 1.f), make_float2(0.5f, m)};
        for (int r = 0; r < reps; ++r) {
            Dft<16, float2>::run(x);
#pragma unroll
            for (int q = 1; q < 16; ++q) x[q] = cmul(x[q], 0.8f, 0.6f);
        }
        float s = 0.f;
#pragma unroll
        for (int m = 0; m < 16; ++m) s += x[m].re.x + x[m].im.y;
        if (s == 12345.678f) sink[0] = s;
    }
    __syncthreads();
    unsigned long long t1 = clock64();
    if (tid == 0) out[blockIdx.x] = t1 - t0;
    if (buf[tid].x == 12345.678f) sink[1] = buf[tid].x;
}

int main() {
    unsigned long long *d;
    float *sink;
    float2 *twg;
    cudaMalloc(&d, 148 * 8);
    cudaMalloc(&sink, 16);
    cudaMalloc(&twg, PL::TWN * 8);
    cudaMemset(twg, 0, PL::TWN * 8);
    const size_t smem = SM::data_bytes + PL::TWN * 8 + 1024;
    const int reps = 200;
    const char *names[] = {"full pass, warps free-running", "shared-memory traffic only", "radix-16 butterfly + twiddle multiply only", "full pass + CTA barrier per pass (lock step)", "two 8-warp groups, own barriers, offset 1300 clk", "two 8-warp groups, own barriers, no offset"};
    for (int mode = 0; mode < 6; ++mode) {
        auto fn = mode == 0 ? k<0> : mode == 1 ? k<1> : mode == 2 ? k<2> : mode == 3 ? k<3> : mode == 4 ? k<4> : k<5>;
        cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        for (int it = 0; it < 2; ++it) { fn<<<148, NT, smem>>>(reps, twg, d, sink); cudaDeviceSynchronize(); }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<unsigned long long> h(148);
        cudaMemcpy(h.data(), d, 148 * 8, cudaMemcpyDeviceToHost);
        double avg = 0;
        for (auto v : h) avg += v;
        avg /= 148.0 * reps;
        printf("%-48s %7.0f cycles per pass per SM (%.2f us @1.965 GHz)\n", names[mode], avg, avg / 1965.0);
    }
    for (int v = 0; v < 8; ++v) {
        const int sync = v & 1;
        auto fn = v == 0 ? kdual<0, 0> : v == 1 ? kdual<1, 0> : v == 2 ? kdual<0, 1> : v == 3 ? kdual<1, 1> : v == 4 ? kdual<0, 2> : v == 5 ? kdual<1, 2> : v == 6 ? kdual<0, 3> : kdual<1, 3>;
        cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        for (int it = 0; it < 2; ++it) { fn<<<148, 256, smem>>>(reps, twg, d, sink); cudaDeviceSynchronize(); }
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<unsigned long long> h(148);
        cudaMemcpy(h.data(), d, 148 * 8, cudaMemcpyDeviceToHost);
        double avg = 0;
        for (auto v : h) avg += v;
        avg /= 148.0 * reps;
        printf("[early=%d] %-48s %7.0f cycles per pass per SM (%.2f us @1.965 GHz)\n", v >> 1,
               sync ? "256 threads, both columns per thread, CTA barrier" : "256 threads, both columns per thread, free-running", avg, avg / 1965.0);
    }
    return 0;
}
