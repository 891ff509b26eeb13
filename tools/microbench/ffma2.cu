// Microbenchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2, sm_100+) throughput per SM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu ; run on a B200.
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

template <int MODE>
__global__ void __launch_bounds__(256) kern(float* out, int iters, float s, float t) {
    float acc[16];
#pragma unroll
    for (int k = 0; k < 16; k++) acc[k] = threadIdx.x * 0.001f + k;
    if (MODE == 0) {            // 16 scalar FMA chains
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int k = 0; k < 16; k++) acc[k] = __fmaf_rn(acc[k], s, t);
        }
    } else if (MODE == 1) {     // 8 packed chains (same flops)
        u64 p[8];
        const u64 ss = pack(s, s), tt = pack(t, t);
#pragma unroll
        for (int k = 0; k < 8; k++) p[k] = pack(acc[2 * k], acc[2 * k + 1]);
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int k = 0; k < 8; k++) p[k] = fma2(p[k], ss, tt);
        }
#pragma unroll
        for (int k = 0; k < 8; k++) unpack(p[k], acc[2 * k], acc[2 * k + 1]);
    } else if (MODE == 2) {     // scalar FMA interleaved with ALU-pipe min (same count each)
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int k = 0; k < 8; k++) { acc[k] = __fmaf_rn(acc[k], s, t); acc[8 + k] = fminf(acc[8 + k], acc[k]); }
        }
    } else if (MODE == 3) {     // packed FMA2 interleaved with ALU-pipe min
        u64 p[4];
        const u64 ss = pack(s, s), tt = pack(t, t);
#pragma unroll
        for (int k = 0; k < 4; k++) p[k] = pack(acc[2 * k], acc[2 * k + 1]);
        for (int i = 0; i < iters; i++) {
#pragma unroll
            for (int k = 0; k < 4; k++) {
                p[k] = fma2(p[k], ss, tt);
                float a, b; unpack(p[k], a, b);
                acc[8 + 2 * k] = fminf(acc[8 + 2 * k], a); acc[9 + 2 * k] = fminf(acc[9 + 2 * k], b);
            }
        }
#pragma unroll
        for (int k = 0; k < 4; k++) unpack(p[k], acc[2 * k], acc[2 * k + 1]);
    }
    float r = 0.f;
#pragma unroll
    for (int k = 0; k < 16; k++) r += acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE>
float run(float* out, int blocks, int iters) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    kern<MODE><<<blocks, 256>>>(out, iters, 0.999f, 0.001f);
    cudaEventRecord(a);
    kern<MODE><<<blocks, 256>>>(out, iters, 0.999f, 0.001f);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 8, iters = 20000;
    float* out; cudaMalloc(&out, (size_t)blocks * 256 * 4);
    const double threads = (double)blocks * 256;
    float t0 = run<0>(out, blocks, iters), t1 = run<1>(out, blocks, iters), t2 = run<2>(out, blocks, iters), t3 = run<3>(out, blocks, iters);
    printf("SMs %d\n", sms);
    printf("scalar FFMA x16       : %.3f ms  %.1f GFMA/s  (%.1f FMA/clk/SM @1.965GHz)\n", t0, threads * 16 * iters / t0 / 1e6, threads * 16 * iters / t0 / 1e6 / sms / 1.965);
    printf("packed FFMA2 x8       : %.3f ms  %.1f GFMA/s  (%.1f FMA/clk/SM)\n", t1, threads * 16 * iters / t1 / 1e6, threads * 16 * iters / t1 / 1e6 / sms / 1.965);
    printf("FFMA x8 + FMNMX x8    : %.3f ms  %.1f Ginst-lanes/s\n", t2, threads * 16 * iters / t2 / 1e6);
    printf("FFMA2 x4 + FMNMX x8   : %.3f ms  (8 FMA + 8 MNMX per iter)\n", t3);
    return 0;
}
