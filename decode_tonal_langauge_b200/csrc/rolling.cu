// K7 trailing-window z-score (ref: preprocess/signal/rolling_zscore.py:28-49, pandas rolling
// mean / std with min_periods=1, ddof=1).
//
//   z[t] = (x[t] - mean_w(t)) / std_w(t),  window = x[max(0, t+1-W) .. t],  n = window length,
//   std_w^2 = (S2 - S1^2 / n) / (n - 1)  ->  sample 0 is NaN (n = 1).
//
// Window sums are differences of float64 prefix sums of d = x - shift (shift = row mean, from
// ecog_row_stats) and d^2.  The prefixes are materialised only at 16-sample granularity
// (P16: 16 B per 16 samples = 1 B per sample of extra traffic):
//   1. rz_block_sums : per 16-sample block sum(d), sum(d^2); per 4096-sample segment totals;
//   2. rz_prefix     : exclusive block scan INSIDE each segment -> P16 (segment-local, so every
//                      quantity that is later differenced has window-scale magnitude, like
//                      pandas' add/remove update, not record-scale magnitude);
//   3. rz_apply      : each thread owns 16 consecutive samples: exact window sums at its first
//                      sample = local prefix + whole-segment totals in between + remainder of the
//                      window's first segment (windows of <= 32 samples are summed directly),
//                      then add / remove in float64 for the next 15 samples.
// HBM-bound: 8 B per sample algorithmic (+ the window tail re-read, which hits L2).
#include "common.cuh"

namespace ecog {

constexpr int kRzThreads = 256;
constexpr int kRzBlk = 16;                          // samples per thread
constexpr int kRzSeg = kRzThreads * kRzBlk;         // samples per CTA

__device__ __forceinline__ void rz_load16(const float* __restrict__ p, int64_t t0, int64_t T, bool vec, float (&v)[kRzBlk]) {
    if (vec && t0 + kRzBlk <= T) {
#pragma unroll
        for (int q = 0; q < kRzBlk / 4; ++q) {
            const float4 f = *reinterpret_cast<const float4*>(p + t0 + 4 * q);
            v[4 * q] = f.x; v[4 * q + 1] = f.y; v[4 * q + 2] = f.z; v[4 * q + 3] = f.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < kRzBlk; ++i) v[i] = (t0 + i >= 0 && t0 + i < T) ? p[t0 + i] : 0.f;
    }
}

// grid (nseg, C).  P16: [C][nblk][2] per-block sums (later overwritten by exclusive prefixes);
// seg: [C][nseg][2] segment totals.
__global__ void __launch_bounds__(kRzThreads)
rz_block_sums(const float* __restrict__ x, int64_t T, int64_t ld, const double* __restrict__ shift,
              double* __restrict__ P16, double* __restrict__ seg, int64_t nblk, int nseg, bool vec) {
    const int64_t row = blockIdx.y;
    const int64_t blk = (int64_t)blockIdx.x * kRzThreads + threadIdx.x;
    const int64_t t0 = blk * kRzBlk;
    const double sh = shift[row];
    double s1 = 0.0, s2 = 0.0;
    if (t0 < T) {
        float v[kRzBlk];
        rz_load16(x + row * ld, t0, T, vec, v);
#pragma unroll
        for (int i = 0; i < kRzBlk; ++i) {
            if (t0 + i < T) { const double d = (double)v[i] - sh; s1 += d; s2 = fma(d, d, s2); }
        }
        P16[(row * nblk + blk) * 2 + 0] = s1;
        P16[(row * nblk + blk) * 2 + 1] = s2;
    }
    __shared__ double w1[kRzThreads / 32], w2[kRzThreads / 32];
    const double a1 = warp_sum(s1), a2 = warp_sum(s2);
    if ((threadIdx.x & 31) == 0) { w1[threadIdx.x >> 5] = a1; w2[threadIdx.x >> 5] = a2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double b1 = 0.0, b2 = 0.0;
#pragma unroll
        for (int k = 0; k < kRzThreads / 32; ++k) { b1 += w1[k]; b2 += w2[k]; }
        seg[(row * nseg + blockIdx.x) * 2 + 0] = b1;
        seg[(row * nseg + blockIdx.x) * 2 + 1] = b2;
    }
}

// exclusive scan of the block sums inside each segment (fixed order: warp shuffles, then 8 warps)
__global__ void __launch_bounds__(kRzThreads)
rz_prefix(double* __restrict__ P16, int64_t nblk) {
    const int64_t row = blockIdx.y;
    const int64_t blk = (int64_t)blockIdx.x * kRzThreads + threadIdx.x;
    const bool ok = blk < nblk;
    double s1 = ok ? P16[(row * nblk + blk) * 2 + 0] : 0.0;
    double s2 = ok ? P16[(row * nblk + blk) * 2 + 1] : 0.0;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double i1 = s1, i2 = s2;                         // inclusive scan within the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double u1 = __shfl_up_sync(0xffffffffu, i1, o), u2 = __shfl_up_sync(0xffffffffu, i2, o);
        if (lane >= o) { i1 += u1; i2 += u2; }
    }
    __shared__ double w1[kRzThreads / 32], w2[kRzThreads / 32];
    if (lane == 31) { w1[warp] = i1; w2[warp] = i2; }
    __syncthreads();
    double o1 = 0.0, o2 = 0.0;
    for (int k = 0; k < warp; ++k) { o1 += w1[k]; o2 += w2[k]; }
    if (ok) {
        P16[(row * nblk + blk) * 2 + 0] = o1 + (i1 - s1);
        P16[(row * nblk + blk) * 2 + 1] = o2 + (i2 - s2);
    }
}

__global__ void __launch_bounds__(kRzThreads)
rz_apply(const float* __restrict__ x, float* __restrict__ y, int64_t T, int64_t ldx, int64_t ldy,
         const double* __restrict__ shift, const double* __restrict__ P16, const double* __restrict__ seg,
         int64_t nblk, int nseg, int64_t W, int nan_to_zero, bool vec) {
    const int64_t row = blockIdx.y;
    const int64_t blk = (int64_t)blockIdx.x * kRzThreads + threadIdx.x;
    const int64_t t0 = blk * kRzBlk;
    if (t0 >= T) return;
    const float* xr = x + row * ldx;
    const double sh = shift[row];
    const double* P = P16 + row * nblk * 2;
    float v[kRzBlk], old[kRzBlk];
    rz_load16(xr, t0, T, vec, v);
    // samples leaving the window while this thread advances: x[t0 + i - W], i = 1..15 (index 0 unused)
    {
        const int64_t o0 = t0 - W;
        if (o0 >= 0 && (o0 & 3) == 0) rz_load16(xr, o0, T, vec, old);
        else {
#pragma unroll
            for (int i = 0; i < kRzBlk; ++i) old[i] = (o0 + i >= 0 && o0 + i < T) ? xr[o0 + i] : 0.f;
        }
    }
    // exact sums over the window [lo0, t0] of the first sample
    const int64_t lo0 = t0 + 1 - W > 0 ? t0 + 1 - W : 0;
    const double d0 = (double)v[0] - sh;
    double S1, S2;
    if (W <= 2 * kRzBlk) {                                   // short window: sum it directly
        S1 = d0; S2 = d0 * d0;
        for (int64_t t = lo0; t < t0; ++t) { const double d = (double)xr[t] - sh; S1 += d; S2 = fma(d, d, S2); }
    } else {
        const double* G = seg + row * nseg * 2;
        const int64_t lb = lo0 / kRzBlk;
        const int sa = (int)(blk / kRzThreads), sl = (int)(lb / kRzThreads);
        double q1 = P[2 * lb], q2 = P[2 * lb + 1];           // segment-local prefix at the window start
        for (int64_t t = lb * kRzBlk; t < lo0; ++t) { const double d = (double)xr[t] - sh; q1 += d; q2 = fma(d, d, q2); }
        S1 = P[2 * blk] + d0; S2 = fma(d0, d0, P[2 * blk + 1]);
        if (sl == sa) { S1 -= q1; S2 -= q2; }
        else {
            double m1 = G[2 * sl] - q1, m2 = G[2 * sl + 1] - q2;      // rest of the window's first segment
            for (int sgm = sl + 1; sgm < sa; ++sgm) { m1 += G[2 * sgm]; m2 += G[2 * sgm + 1]; }
            S1 += m1; S2 += m2;
        }
    }
    float out[kRzBlk];
#pragma unroll
    for (int i = 0; i < kRzBlk; ++i) {
        const int64_t t = t0 + i;
        const double d = (double)v[i] - sh;
        if (i > 0) {
            S1 += d; S2 = fma(d, d, S2);
            if (t - W >= 0) { const double r = (double)old[i] - sh; S1 -= r; S2 = fma(-r, r, S2); }
        }
        const int64_t lo = t + 1 - W > 0 ? t + 1 - W : 0;
        const double n = (double)(t + 1 - lo);
        const double mean = S1 / n;
        double var = (S2 - S1 * S1 / n) / (n - 1.0);       // n == 1: 0 / 0 = NaN, like pandas' ddof=1
        if (var < 0.0) var = 0.0;
        float z = (float)((d - mean) / sqrt(var));
        if (nan_to_zero && isnan(z)) z = 0.f;
        out[i] = z;
    }
    float* yr = y + row * ldy + t0;
    if (vec && t0 + kRzBlk <= T) {
#pragma unroll
        for (int q = 0; q < kRzBlk / 4; ++q)
            *reinterpret_cast<float4*>(yr + 4 * q) = make_float4(out[4 * q], out[4 * q + 1], out[4 * q + 2], out[4 * q + 3]);
    } else {
#pragma unroll
        for (int i = 0; i < kRzBlk; ++i)
            if (t0 + i < T) yr[i] = out[i];
    }
}

static size_t rz_align(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace ecog

using namespace ecog;

extern "C" size_t ecog_rolling_workspace(int64_t C, int64_t T) {
    const int64_t nblk = ceil_div(T, kRzBlk), nseg = ceil_div(T, kRzSeg);
    return rz_align((size_t)C * nblk * 2 * sizeof(double)) + rz_align((size_t)C * nseg * 2 * sizeof(double));
}

extern "C" int ecog_rolling_zscore(const float* d_x, float* d_y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                                   int64_t window, const double* d_shift, int nan_to_zero,
                                   void* d_workspace, size_t workspace_bytes, ecog_stream_t stream) {
    if (C <= 0 || C > 65535 || T <= 0 || ldx < T || ldy < T) return fail(ECOG_E_VALUE, "ecog_rolling_zscore: bad shape");
    if (window <= 1) return fail(ECOG_E_VALUE, "window_size must be greater than 1.");
    if (!d_shift) return fail(ECOG_E_VALUE, "ecog_rolling_zscore: per-row shift is required");
    if (workspace_bytes < ecog_rolling_workspace(C, T))
        return fail(ECOG_E_WORKSPACE, "ecog_rolling_zscore: workspace %zu < %zu", workspace_bytes, ecog_rolling_workspace(C, T));
    const int64_t nblk = ceil_div(T, kRzBlk);
    const int nseg = (int)ceil_div(T, kRzSeg);
    double* P16 = (double*)d_workspace;
    double* seg = (double*)((char*)d_workspace + rz_align((size_t)C * nblk * 2 * sizeof(double)));
    const bool vec = aligned16(d_x) && aligned16(d_y) && ldx % 4 == 0 && ldy % 4 == 0;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)nseg, (unsigned)C);
    rz_block_sums<<<grid, kRzThreads, 0, st>>>(d_x, T, ldx, d_shift, P16, seg, nblk, nseg, vec);
    ECOG_TRY(check_launch("rz_block_sums"));
    rz_prefix<<<grid, kRzThreads, 0, st>>>(P16, nblk);
    ECOG_TRY(check_launch("rz_prefix"));
    rz_apply<<<grid, kRzThreads, 0, st>>>(d_x, d_y, T, ldx, ldy, d_shift, P16, seg, nblk, nseg, window, nan_to_zero, vec);
    return check_launch("rz_apply");
}
