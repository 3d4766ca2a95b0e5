// Microbenchmark: FP64 issue rate and latency on sm_100a (B200), alone and next to FP32 work.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_pipes fp64_pipes.cu && ./fp64_pipes
#include <cstdio>
#include <cuda_runtime.h>

// MODE 0: DFMA a = a*m + c (two shared operands)      1: DFMA with three distinct registers
//      2: DFMA + FFMA interleaved one to one (independent chains)   3: F2F f32->f64->f32 round trip
//      4: DADD
// CH independent chains per thread (CH = 1: the dependent-issue latency)
template <int MODE, int CH>
__global__ void __launch_bounds__(256) bench(double* out, double s, int iters) {
    double a[CH];
    float f[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) { a[i] = threadIdx.x * 0.001 + i; f[i] = (float)a[i]; }
    const double m = s, c = s * 0.5;
    const float mf = (float)s, cf = (float)c;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (MODE == 0) a[i] = fma(a[i], m, c);
            if (MODE == 1) a[i] = fma(a[i], a[(i + 1) % CH], a[(i + 2) % CH]);
            if (MODE == 2) { a[i] = fma(a[i], m, c); f[i] = fmaf(f[i], mf, cf); }
            if (MODE == 3) { a[i] = (double)f[i] + 0.0; f[i] = (float)a[i]; }
            if (MODE == 4) a[i] = a[i] + c;
        }
    }
    double r = 0.0;
#pragma unroll
    for (int i = 0; i < CH; ++i) r += a[i] + (double)f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int MODE, int CH> void run(const char* name, double* d, int ctas_per_sm, double per_iter) {
    const int iters = 4096, grid = 148 * ctas_per_sm;
    bench<MODE, CH><<<grid, 256>>>(d, 1.0000001, 16);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<MODE, CH><<<grid, 256>>>(d, 1.0000001, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double winstr = (double)grid * 8 * iters * CH * per_iter;          // warp instructions of the named kind
    const double per_clk_sm = winstr / (ms * 1e-3) / (clk * 1e3) / 148.0;
    const double cyc_per_iter = (ms * 1e-3) * (clk * 1e3) / iters;           // cycles per loop iteration of one thread
    printf("%-44s %d warps/SM  %.3f ms  %.3f warp-instr/clk/SM  %.1f cycles per iteration\n", name, 8 * ctas_per_sm, ms,
           per_clk_sm, cyc_per_iter);
}
int main() {
    double* d; cudaMalloc(&d, 148 * 8 * 256 * sizeof(double));
    for (int w : {1, 2, 4}) {
        run<0, 1>("DFMA dependent chain (latency)", d, w, 1);
        run<0, 4>("DFMA 4 chains, 2 shared operands", d, w, 1);
        run<0, 8>("DFMA 8 chains, 2 shared operands", d, w, 1);
        run<1, 8>("DFMA 8 chains, 3 distinct registers", d, w, 1);
        run<4, 8>("DADD 8 chains", d, w, 1);
        run<2, 8>("DFMA + FFMA one to one, 8 chains (DFMA rate)", d, w, 1);
        run<3, 8>("F2F f32->f64, DADD, F2F f64->f32 (F2F rate)", d, w, 2);
    }
    return 0;
}
