// K1 common-average re-reference and K2 per-channel statistics / z-score.
//
// HBM-bound streaming kernels.  Layout: float32 (C, T) row-major, row stride ld.
//   car_fused      : one read + one write per sample (8 B).  A CTA stages a
//                    [C x TT] column strip in shared memory with cp.async, reduces the
//                    columns there, subtracts and streams the strip back out.
//   car_colsum/apply : two-phase form for channel-sharded recordings (the caller
//                    all-reduces the T column sums in between).
//   row_stats      : float64 sum / sum-of-squares about the row's first sample,
//                    fixed-order two-level reduction (deterministic).
//   zscore_apply   : (x - mean) / std in float32, 128-bit accesses.
#include "common.cuh"

namespace ecog {

thread_local char g_err[512] = "";
thread_local int64_t g_launches = 0;
thread_local char g_launch_log[4096] = "";
thread_local int g_launch_log_len = 0;

constexpr int kCarThreads = 256;

template <int TT, bool VEC>
__global__ void __launch_bounds__(kCarThreads)
car_fused_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int64_t T, int64_t ld, int64_t ldy,
                 const float* __restrict__ w, float inv_count) {
    extern __shared__ __align__(16) float smem[];
    constexpr int V = TT / 4;                 // float4 column groups
    constexpr int RG = kCarThreads / V;       // row groups reducing in parallel
    float* tile = smem;                       // [C][TT]
    float* psum = smem + (size_t)C * TT;      // [RG][TT]
    float* mean = psum + RG * TT;             // [TT]
    const int64_t t0 = (int64_t)blockIdx.x * TT;
    const int tid = threadIdx.x;

    if (VEC) {
        for (int i = tid; i < C * V; i += kCarThreads) {
            int row = i / V, v = i - row * V;
            int64_t t = t0 + 4 * v;
            bool ok = t < T;
            cp_async16_zfill(&tile[row * TT + 4 * v], x + (int64_t)row * ld + (ok ? t : 0), ok);
        }
    } else {
        for (int i = tid; i < C * TT; i += kCarThreads) {
            int row = i / TT, v = i - row * TT;
            int64_t t = t0 + v;
            bool ok = t < T;
            cp_async4_zfill(&tile[row * TT + v], x + (int64_t)row * ld + (ok ? t : 0), ok);
        }
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    {   // column partial sums: thread (cg, rg) adds rows rg, rg+RG, ... of float4 column group cg
        const int cg = tid % V, rg = tid / V;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int row = rg; row < C; row += RG) {
            float4 v = *reinterpret_cast<const float4*>(&tile[row * TT + 4 * cg]);
            float wr = w ? w[row] : 1.0f;
            acc.x = fmaf(wr, v.x, acc.x); acc.y = fmaf(wr, v.y, acc.y);
            acc.z = fmaf(wr, v.z, acc.z); acc.w = fmaf(wr, v.w, acc.w);
        }
        *reinterpret_cast<float4*>(&psum[rg * TT + 4 * cg]) = acc;
    }
    __syncthreads();
    if (tid < TT) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < RG; ++r) s += psum[r * TT + tid];
        mean[tid] = s * inv_count;
    }
    __syncthreads();

    if (VEC) {
        for (int i = tid; i < C * V; i += kCarThreads) {
            int row = i / V, v = i - row * V;
            int64_t t = t0 + 4 * v;
            if (t < T) {
                float4 a = *reinterpret_cast<const float4*>(&tile[row * TT + 4 * v]);
                float4 m = *reinterpret_cast<const float4*>(&mean[4 * v]);
                a.x -= m.x; a.y -= m.y; a.z -= m.z; a.w -= m.w;
                stg_stream(reinterpret_cast<float4*>(y + (int64_t)row * ldy + t), a);
            }
        }
    } else {
        for (int i = tid; i < C * TT; i += kCarThreads) {
            int row = i / TT, v = i - row * TT;
            int64_t t = t0 + v;
            if (t < T) y[(int64_t)row * ldy + t] = tile[row * TT + v] - mean[v];
        }
    }
}

// two-phase: column sums over this shard's rows (float32 accumulate, row order)
__global__ void __launch_bounds__(256)
car_colsum_kernel(const float* __restrict__ x, int C, int64_t T, int64_t ld,
                  const float* __restrict__ w, float* __restrict__ colsum, bool vec) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        int64_t t = 4 * i;
        if (t >= T) return;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 8
        for (int row = 0; row < C; ++row) {
            float4 v = ldg_stream(reinterpret_cast<const float4*>(x + (int64_t)row * ld + t));
            float wr = w ? w[row] : 1.0f;
            acc.x = fmaf(wr, v.x, acc.x); acc.y = fmaf(wr, v.y, acc.y);
            acc.z = fmaf(wr, v.z, acc.z); acc.w = fmaf(wr, v.w, acc.w);
        }
        *reinterpret_cast<float4*>(colsum + t) = acc;
    } else {
        if (i >= T) return;
        float acc = 0.f;
        for (int row = 0; row < C; ++row) acc = fmaf(w ? w[row] : 1.0f, x[(int64_t)row * ld + i], acc);
        colsum[i] = acc;
    }
}

__global__ void __launch_bounds__(256)
car_apply_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t T, int64_t ld, int64_t ldy,
                 const float* __restrict__ colsum, float inv_count, bool vec) {
    const int64_t row = blockIdx.y;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        int64_t t = 4 * i;
        if (t >= T) return;
        float4 v = ldg_stream(reinterpret_cast<const float4*>(x + row * ld + t));
        float4 m = *reinterpret_cast<const float4*>(colsum + t);
        v.x -= m.x * inv_count; v.y -= m.y * inv_count; v.z -= m.z * inv_count; v.w -= m.w * inv_count;
        stg_stream(reinterpret_cast<float4*>(y + row * ldy + t), v);
    } else {
        if (i >= T) return;
        y[row * ldy + i] = x[row * ld + i] - colsum[i] * inv_count;
    }
}

// ------------------------------------------------------------------ statistics
constexpr int kStatThreads = 256;
constexpr int64_t kStatSlab = 64 * 1024;     // samples per CTA

__global__ void __launch_bounds__(kStatThreads)
row_stats_partial_kernel(const float* __restrict__ x, int64_t ld, int64_t t0, int64_t t1,
                         int nslab, double* __restrict__ partial) {
    const int64_t row = blockIdx.y;
    const float* p = x + row * ld;
    const double shift = (double)p[t0];
    int64_t a = t0 + (int64_t)blockIdx.x * kStatSlab;
    int64_t b = a + kStatSlab < t1 ? a + kStatSlab : t1;
    double s1 = 0.0, s2 = 0.0;
    // scalar head to a 16-byte boundary, float4 body, scalar tail
    int64_t i = a + threadIdx.x;
    const bool can_vec = ((reinterpret_cast<uintptr_t>(p) & 15u) == 0);
    if (can_vec) {
        int64_t a4 = (a + 3) & ~int64_t(3), b4 = b & ~int64_t(3);
        if (a4 > b4) { a4 = b; b4 = b; }
        for (int64_t j = a + threadIdx.x; j < a4; j += kStatThreads) {
            double d = (double)p[j] - shift; s1 += d; s2 = fma(d, d, s2);
        }
        for (int64_t j = a4 + 4 * (int64_t)threadIdx.x; j < b4; j += 4 * kStatThreads) {
            float4 v = ldg_stream(reinterpret_cast<const float4*>(p + j));
            double d0 = (double)v.x - shift, d1 = (double)v.y - shift;
            double d2 = (double)v.z - shift, d3 = (double)v.w - shift;
            s1 += (d0 + d1) + (d2 + d3);
            s2 = fma(d0, d0, s2); s2 = fma(d1, d1, s2); s2 = fma(d2, d2, s2); s2 = fma(d3, d3, s2);
        }
        for (int64_t j = b4 + threadIdx.x; j < b; j += kStatThreads) {
            if (j >= a4) { double d = (double)p[j] - shift; s1 += d; s2 = fma(d, d, s2); }
        }
    } else {
        for (; i < b; i += kStatThreads) {
            double d = (double)p[i] - shift; s1 += d; s2 = fma(d, d, s2);
        }
    }
    __shared__ double sh1[kStatThreads / 32], sh2[kStatThreads / 32];
    s1 = warp_sum(s1); s2 = warp_sum(s2);
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sh1[warp] = s1; sh2[warp] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a1 = 0.0, a2 = 0.0;
#pragma unroll
        for (int k = 0; k < kStatThreads / 32; ++k) { a1 += sh1[k]; a2 += sh2[k]; }
        partial[(row * nslab + blockIdx.x) * 2 + 0] = a1;
        partial[(row * nslab + blockIdx.x) * 2 + 1] = a2;
    }
}

__global__ void row_stats_final_kernel(const float* __restrict__ x, int64_t ld, int64_t t0, int64_t t1,
                                       int nslab, const double* __restrict__ partial,
                                       double* __restrict__ mean, double* __restrict__ stdev, int C) {
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= C) return;
    double s1 = 0.0, s2 = 0.0;
    for (int k = 0; k < nslab; ++k) {
        s1 += partial[((int64_t)row * nslab + k) * 2 + 0];
        s2 += partial[((int64_t)row * nslab + k) * 2 + 1];
    }
    double n = (double)(t1 - t0);
    double shift = (double)x[(int64_t)row * ld + t0];
    double m = s1 / n;
    double var = s2 / n - m * m;
    if (var < 0.0) var = 0.0;
    mean[row] = shift + m;
    stdev[row] = sqrt(var);
}

__global__ void __launch_bounds__(256)
zscore_apply_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t T, int64_t ld, int64_t ldy,
                    const double* __restrict__ mean, const double* __restrict__ stdev,
                    int nan_to_zero, bool vec) {
    const int64_t row = blockIdx.y;
    const float m = (float)mean[row], s = (float)stdev[row];
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        int64_t t = 4 * i;
        if (t >= T) return;
        float4 v = ldg_stream(reinterpret_cast<const float4*>(x + row * ld + t));
        v.x = (v.x - m) / s; v.y = (v.y - m) / s; v.z = (v.z - m) / s; v.w = (v.w - m) / s;
        if (nan_to_zero) {
            if (isnan(v.x)) v.x = 0.f; if (isnan(v.y)) v.y = 0.f;
            if (isnan(v.z)) v.z = 0.f; if (isnan(v.w)) v.w = 0.f;
        }
        stg_stream(reinterpret_cast<float4*>(y + row * ldy + t), v);
    } else {
        if (i >= T) return;
        float v = (x[row * ld + i] - m) / s;
        if (nan_to_zero && isnan(v)) v = 0.f;
        y[row * ldy + i] = v;
    }
}

static bool vec_ok(const void* a, const void* b, int64_t T, int64_t ld, int64_t ldy = 0) {
    return aligned16(a) && (b == nullptr || aligned16(b)) && (T % 4 == 0) && (ld % 4 == 0) && (ldy % 4 == 0);
}

template <int TT, bool VEC>
static int launch_car_fused(const float* x, float* y, int C, int64_t T, int64_t ld, int64_t ldy, const float* w,
                            float inv, cudaStream_t st) {
    constexpr int RG = kCarThreads / (TT / 4);
    size_t smem = ((size_t)C * TT + (size_t)RG * TT + TT) * sizeof(float);
    auto k = car_fused_kernel<TT, VEC>;
    ECOG_TRY((smem_attr<car_fused_kernel<TT, VEC>>(smem)));
    k<<<(unsigned)ceil_div(T, TT), kCarThreads, smem, st>>>(x, y, C, T, ld, ldy, w, inv);
    return check_launch("car_fused");
}

}  // namespace ecog

using namespace ecog;

extern "C" int ecog_abi_version(void) { return ECOG_ABI_VERSION; }
extern "C" const char* ecog_last_error(void) { return g_err; }
extern "C" int64_t ecog_launch_count(void) { return g_launches; }
extern "C" const char* ecog_launch_log(int reset) {
    static thread_local char out[4096];
    memcpy(out, g_launch_log, sizeof(out));
    if (reset) { g_launch_log_len = 0; g_launch_log[0] = 0; }
    return out;
}

extern "C" int ecog_car(const float* d_x, float* d_y, int64_t C, int64_t T, int64_t ld, int64_t ldy,
                        const float* d_w, double inv_count, ecog_stream_t stream) {
    if (C <= 0 || T <= 0 || ld < T || ldy < T) return fail(ECOG_E_VALUE, "ecog_car: bad shape C=%lld T=%lld ld=%lld ldy=%lld",
                                                 (long long)C, (long long)T, (long long)ld, (long long)ldy);
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = vec_ok(d_x, d_y, T, ld, ldy);
    const float inv = (float)inv_count;
    // widest strip whose [C x TT] tile leaves room for three CTAs per SM, else one
    const size_t budget3 = 72 * 1024, budget1 = 200 * 1024;
    auto bytes = [&](int tt) { return ((size_t)C * tt + (size_t)(kCarThreads / (tt / 4)) * tt + tt) * 4; };
    int tt = 0;
    for (int cand : {128, 64, 32}) if (bytes(cand) <= budget3) { tt = cand; break; }
    if (!tt) for (int cand : {128, 64, 32}) if (bytes(cand) <= budget1) { tt = cand; break; }
    if (!tt) return fail(ECOG_E_UNSUPPORTED, "ecog_car: C=%lld too large for the fused strip; use colsum/apply",
                         (long long)C);
    if (vec) {
        if (tt == 128) return launch_car_fused<128, true>(d_x, d_y, (int)C, T, ld, ldy, d_w, inv, st);
        if (tt == 64) return launch_car_fused<64, true>(d_x, d_y, (int)C, T, ld, ldy, d_w, inv, st);
        return launch_car_fused<32, true>(d_x, d_y, (int)C, T, ld, ldy, d_w, inv, st);
    }
    if (tt == 128) return launch_car_fused<128, false>(d_x, d_y, (int)C, T, ld, ldy, d_w, inv, st);
    if (tt == 64) return launch_car_fused<64, false>(d_x, d_y, (int)C, T, ld, ldy, d_w, inv, st);
    return launch_car_fused<32, false>(d_x, d_y, (int)C, T, ld, ldy, d_w, inv, st);
}

extern "C" int ecog_car_colsum(const float* d_x, int64_t C, int64_t T, int64_t ld, const float* d_w,
                               float* d_colsum, ecog_stream_t stream) {
    if (C <= 0 || T <= 0 || ld < T) return fail(ECOG_E_VALUE, "ecog_car_colsum: bad shape");
    const bool vec = vec_ok(d_x, d_colsum, T, ld);
    int64_t n = vec ? T / 4 : T;
    car_colsum_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(
        d_x, (int)C, T, ld, d_w, d_colsum, vec);
    return check_launch("car_colsum");
}

extern "C" int ecog_car_apply(const float* d_x, float* d_y, int64_t C, int64_t T, int64_t ld, int64_t ldy,
                              const float* d_colsum, double inv_count, ecog_stream_t stream) {
    if (C <= 0 || T <= 0 || ld < T || ldy < T || C > 65535) return fail(ECOG_E_VALUE, "ecog_car_apply: bad shape");
    const bool vec = vec_ok(d_x, d_y, T, ld, ldy) && aligned16(d_colsum);
    int64_t n = vec ? T / 4 : T;
    dim3 grid((unsigned)ceil_div(n, 256), (unsigned)C);
    car_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_x, d_y, T, ld, ldy, d_colsum, (float)inv_count, vec);
    return check_launch("car_apply");
}

static int stat_slabs(int64_t n) { return (int)ceil_div(n, kStatSlab); }

extern "C" size_t ecog_row_stats_workspace(int64_t C, int64_t T) {
    return (size_t)C * stat_slabs(T) * 2 * sizeof(double);
}

extern "C" int ecog_row_stats(const float* d_x, int64_t C, int64_t T, int64_t ld, int64_t t0, int64_t t1,
                              double* d_mean, double* d_std, void* d_workspace, size_t workspace_bytes,
                              ecog_stream_t stream) {
    if (C <= 0 || T <= 0 || ld < T || C > 65535) return fail(ECOG_E_VALUE, "ecog_row_stats: bad shape");
    if (t0 < 0 || t1 > T) return fail(ECOG_E_VALUE, "Reference time indices are out of bounds.");
    if (t0 >= t1) return fail(ECOG_E_VALUE, "Start time must be less than end time.");
    int nslab = stat_slabs(t1 - t0);
    if (workspace_bytes < (size_t)C * nslab * 2 * sizeof(double))
        return fail(ECOG_E_WORKSPACE, "ecog_row_stats: workspace %zu too small", workspace_bytes);
    cudaStream_t st = (cudaStream_t)stream;
    double* partial = (double*)d_workspace;
    dim3 grid((unsigned)nslab, (unsigned)C);
    row_stats_partial_kernel<<<grid, kStatThreads, 0, st>>>(d_x, ld, t0, t1, nslab, partial);
    ECOG_TRY(check_launch("row_stats_partial"));
    row_stats_final_kernel<<<(unsigned)ceil_div(C, 128), 128, 0, st>>>(d_x, ld, t0, t1, nslab, partial,
                                                                        d_mean, d_std, (int)C);
    return check_launch("row_stats_final");
}

extern "C" int ecog_zscore_apply(const float* d_x, float* d_y, int64_t C, int64_t T, int64_t ld, int64_t ldy,
                                 const double* d_mean, const double* d_std, int nan_to_zero,
                                 ecog_stream_t stream) {
    if (C <= 0 || T <= 0 || ld < T || ldy < T || C > 65535) return fail(ECOG_E_VALUE, "ecog_zscore_apply: bad shape");
    const bool vec = vec_ok(d_x, d_y, T, ld, ldy);
    int64_t n = vec ? T / 4 : T;
    dim3 grid((unsigned)ceil_div(n, 256), (unsigned)C);
    zscore_apply_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_x, d_y, T, ld, ldy, d_mean, d_std, nan_to_zero, vec);
    return check_launch("zscore_apply");
}

extern "C" int ecog_copy2d(const float* d_src, int64_t ld_src, float* d_dst, int64_t ld_dst, int64_t rows,
                           int64_t cols, ecog_stream_t stream) {
    if (rows <= 0 || cols <= 0 || ld_src < cols || ld_dst < cols) return fail(ECOG_E_VALUE, "ecog_copy2d: bad shape");
    ECOG_CUDA(cudaMemcpy2DAsync(d_dst, (size_t)ld_dst * sizeof(float), d_src, (size_t)ld_src * sizeof(float),
                                (size_t)cols * sizeof(float), (size_t)rows, cudaMemcpyDeviceToDevice,
                                (cudaStream_t)stream));
    return ECOG_OK;
}
