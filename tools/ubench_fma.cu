// Micro-benchmark: issue cost of FFMA vs FFMA2 (fma.rn.f32x2) and their mix with ALU ops on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ubench_fma ubench_fma.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long ffma2s(float a, unsigned long long b, unsigned long long c) {
    unsigned long long aa, d;
    asm volatile("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(aa), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ float ffma(float a, float b, float c) {
    float d;
    asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}
__device__ __forceinline__ unsigned iadd(unsigned a, unsigned b) {
    unsigned d;
    asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ unsigned prmt(unsigned a, unsigned b, unsigned c) {
    unsigned d;
    asm volatile("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

template <int MODE>
__global__ void bench(float *out, int iters, float seed) {
    float f[8];
    unsigned long long p[8];
    unsigned u[8];
    for (int i = 0; i < 8; ++i) {
        f[i] = seed + i + threadIdx.x;
        p[i] = ((unsigned long long) __float_as_uint(seed + i)) << 32 | __float_as_uint(seed * i);
        u[i] = threadIdx.x + i;
    }
    const float m = seed * 0.5f;
    const unsigned long long mm = ((unsigned long long) __float_as_uint(m)) << 32 | __float_as_uint(m);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) f[i] = ffma(f[i], m, m);                       // 8 FFMA
            if (MODE == 1) p[i] = ffma2(p[i], mm, mm);                    // 8 FFMA2
            if (MODE == 2) { f[i] = ffma(f[i], m, m); u[i] = iadd(u[i], u[(i + 1) & 7]); }      // 8 FFMA + 8 IADD
            if (MODE == 3) { p[i] = ffma2(p[i], mm, mm); u[i] = iadd(u[i], u[(i + 1) & 7]); }   // 8 FFMA2 + 8 IADD
            if (MODE == 4) { p[i] = ffma2(p[i], mm, mm); u[i] = iadd(u[i], u[(i + 1) & 7]); u[i] = prmt(u[i], u[(i + 3) & 7], 0x5410); }  // 8 FFMA2 + 16 ALU
            if (MODE == 5) p[i] = ffma2s(m, p[i], mm);                    // 8 FFMA2 with scalar broadcast operand
            if (MODE == 6) { f[i] = ffma(f[i], m, m); f[i] = ffma(f[i], m, seed); u[i] = iadd(u[i], u[(i + 1) & 7]); }  // 16 FFMA + 8 IADD
            if (MODE == 7) u[i] = iadd(u[i], u[(i + 1) & 7]);            // 8 IADD
        }
    }
    float acc = 0;
    for (int i = 0; i < 8; ++i) acc += f[i] + (float) (p[i] & 0xFFFF) + (float) u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int MODE>
void run(const char *name, int ops_per_iter, float *out, int sms) {
    const int iters = 1 << 14, threads = 256, blocks = sms * 8;
    bench<MODE><<<blocks, threads>>>(out, 16, 1.0f);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    bench<MODE><<<blocks, threads>>>(out, iters, 1.0f);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    const double warp_insts = (double) blocks * threads / 32 * iters * ops_per_iter;
    // per SM sub-partition (4 per SM) instruction rate, assuming 1.965 GHz
    const double per_smsp_per_clk = warp_insts / (ms * 1e-3) / (sms * 4.0) / 1.965e9;
    printf("%-34s %8.3f ms  %6.3f warp-inst/clk/SMSP (@1.965GHz)  err=%s\n", name, ms, per_smsp_per_clk,
           cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    float *out;
    cudaMalloc(&out, (size_t) p.multiProcessorCount * 8 * 256 * sizeof(float));
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
    run<0>("8 FFMA", 8, out, p.multiProcessorCount);
    run<1>("8 FFMA2", 8, out, p.multiProcessorCount);
    run<5>("8 FFMA2 (scalar bcast operand)", 8, out, p.multiProcessorCount);
    run<7>("8 IADD", 8, out, p.multiProcessorCount);
    run<2>("8 FFMA + 8 IADD", 16, out, p.multiProcessorCount);
    run<3>("8 FFMA2 + 8 IADD", 16, out, p.multiProcessorCount);
    run<4>("8 FFMA2 + 8 IADD + 8 PRMT", 24, out, p.multiProcessorCount);
    run<6>("16 FFMA + 8 IADD", 24, out, p.multiProcessorCount);
    return 0;
}
