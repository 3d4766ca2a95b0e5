// Microbenchmark: copy bandwidth of the IIR kernel's access pattern -- 75 776 concurrent streams
// (one per chunk-thread), each advancing SEG bytes per stage, against a plain linear copy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o strided_streams strided_streams.cu && ./strided_streams
#include <cstdio>
#include <cuda_runtime.h>

template <int PIECES, int DEPTH>   // PIECES x 16 B per stream and stage, DEPTH stages in flight per thread
__global__ void __launch_bounds__(256, 2)
streams(const float4* __restrict__ in, float4* __restrict__ out, long long L4, int nStages) {
    // 256 streams per CTA; thread t moves piece t % PIECES of streams t / PIECES + j * (256 / PIECES)
    const int tid = threadIdx.x, pc = tid % PIECES;
    const long long base = (long long)blockIdx.x * 256;
    for (int st = 0; st < nStages; st += DEPTH) {
        float4 v[DEPTH][PIECES];
#pragma unroll
        for (int d = 0; d < DEPTH; ++d)
#pragma unroll
            for (int j = 0; j < PIECES; ++j) {
                const long long r = base + tid / PIECES + j * (256 / PIECES);
                v[d][j] = in[r * L4 + (long long)(st + d) * PIECES + pc];
            }
#pragma unroll
        for (int d = 0; d < DEPTH; ++d)
#pragma unroll
            for (int j = 0; j < PIECES; ++j) {
                const long long r = base + tid / PIECES + j * (256 / PIECES);
                out[r * L4 + (long long)(st + d) * PIECES + pc] = v[d][j];
            }
    }
}

__global__ void linear(const float4* __restrict__ in, float4* __restrict__ out, long long n) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) out[i] = in[i];
}

template <typename F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaEventRecord(a); for (int i = 0; i < 5; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); return ms / 5;
}

int main() {
    const long long L = 24336, nStreams = 296LL * 256, n = nStreams * L;     // floats: 7.38 GB
    float4 *in, *out; cudaMalloc(&in, n * 4); cudaMalloc(&out, n * 4); cudaMemset(in, 1, n * 4);
    const double gb = 2.0 * n * 4 / 1e9;
    float ms = timeit([&] { linear<<<148 * 16, 256>>>(in, out, n / 4); });
    printf("linear copy                         %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
    ms = timeit([&] { streams<4, 4><<<296, 256>>>(in, out, L / 4, (int)(L / 16)); });
    printf("75776 streams x  64 B, 4 in flight  %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
    ms = timeit([&] { streams<4, 8><<<296, 256>>>(in, out, L / 4, (int)(L / 16)); });
    printf("75776 streams x  64 B, 8 in flight  %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
    ms = timeit([&] { streams<8, 4><<<296, 256>>>(in, out, L / 4, (int)(L / 32)); });
    printf("75776 streams x 128 B, 4 in flight  %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
    ms = timeit([&] { streams<16, 2><<<296, 256>>>(in, out, L / 4, (int)(L / 64)); });
    printf("75776 streams x 256 B, 2 in flight  %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
    // fewer, longer streams (same bytes)
    ms = timeit([&] { streams<4, 8><<<148, 256>>>(in, out, 2 * L / 4, (int)(2 * L / 16)); });
    printf("37888 streams x  64 B, 8 in flight  %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
    ms = timeit([&] { streams<4, 8><<<74, 256>>>(in, out, 4 * L / 4, (int)(4 * L / 16)); });
    printf("18944 streams x  64 B, 8 in flight  %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
    ms = timeit([&] { streams<4, 4><<<592, 256>>>(in, out, L / 2 / 4, (int)(L / 2 / 16)); });
    printf("151552 streams x 64 B, 4 in flight  %.3f ms  %.0f GB/s\n", ms, gb / ms * 1e3);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
