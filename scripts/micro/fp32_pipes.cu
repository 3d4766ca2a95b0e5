// Microbenchmark: issue rate of scalar vs packed FP32 on sm_100a (B200).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_pipes fp32_pipes.cu && ./fp32_pipes
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }

template <int MODE>
__global__ void __launch_bounds__(256) bench(float* out, float s, int iters) {
    float a[16];
    u64 p[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { a[i] = threadIdx.x * 0.001f + i; p[i] = pk(a[i], a[i] + 0.5f); }
    const float m = s, c = s * 0.5f;
    const u64 mm = pk(m, m * 1.0001f), cc = pk(c, c * 0.999f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (MODE == 0) a[i] = fmaf(a[i], m, c);                                   // FFMA R,R,R,R
            if (MODE == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(mm), "l"(cc));
            if (MODE == 2) a[i] = a[i] + c;                                           // FADD
            if (MODE == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(cc));
            if (MODE == 4) a[i] = a[i] * m;                                           // FMUL
            if (MODE == 5) a[i] = fmaf(a[i], a[(i + 1) & 15], a[(i + 2) & 15]);       // FFMA, 3 distinct regs
            if (MODE == 6) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(p[(i + 1) & 15]), "l"(p[(i + 2) & 15]));
        }
    }
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += a[i] + lo(p[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE> void run(const char* name, float* d) {
    const int iters = 8192, grid = 148 * 8;
    bench<MODE><<<grid, 256>>>(d, 1.0001f, 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<MODE><<<grid, 256>>>(d, 1.0001f, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double winstr = (double)grid * 8 * iters * 16;          // warp instructions
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double per_clk_sm = winstr / (ms * 1e-3) / (clk * 1e3) / 148.0;
    printf("%-28s %.3f ms  %.2f warp-instr/clk/SM (at %d MHz nominal)\n", name, ms, per_clk_sm, clk / 1000);
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * sizeof(float));
    run<0>("FFMA  r,r,R,R (2 shared)", d);
    run<5>("FFMA  3 distinct regs", d);
    run<1>("FFMA2 (2 shared)", d);
    run<6>("FFMA2 3 distinct regs", d);
    run<2>("FADD", d);
    run<3>("FADD2", d);
    run<4>("FMUL", d);
    return 0;
}
