// K9 per-(channel, timepoint) one-way ANOVA and K10 longest significant run.
//
// anova_f: one thread per (c, t); the event loop streams the (N, C, L) epoch tensor
// with stride C*L (coalesced along t), float64 accumulators about the first event's
// value (the F statistic is shift invariant; scipy centres on the grand mean for the
// same reason).  4 B read per epoch element -> HBM-bound.  p = fdtrc(G-1, N-G, F) is
// evaluated on the device with the continued fraction of the regularised incomplete
// beta function (modified Lentz), so no host pass over the (C, L) result is needed.
// sig_runlength: one warp per channel, ballot of p < threshold, longest run of ones.
#include "common.cuh"

namespace ecog {

constexpr int kAnovaThreads = 128;
constexpr int kLabelChunk = 2048;

// I_x(a, b), xc = 1 - x supplied by the caller to avoid cancellation
__device__ double betacf(double a, double b, double x) {
    const double FPMIN = 1e-300, EPS = 1e-16;
    double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (fabs(d) < FPMIN) d = FPMIN;
    d = 1.0 / d;
    double h = d;
    for (int m = 1; m <= 5000; ++m) {
        double m2 = 2.0 * m;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d; if (fabs(d) < FPMIN) d = FPMIN;
        c = 1.0 + aa / c; if (fabs(c) < FPMIN) c = FPMIN;
        d = 1.0 / d;
        h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d; if (fabs(d) < FPMIN) d = FPMIN;
        c = 1.0 + aa / c; if (fabs(c) < FPMIN) c = FPMIN;
        d = 1.0 / d;
        double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < EPS) break;
    }
    return h;
}

__device__ double incbeta(double a, double b, double x, double xc) {
    if (x <= 0.0) return 0.0;
    if (xc <= 0.0) return 1.0;
    double lbeta = lgamma(a) + lgamma(b) - lgamma(a + b);
    double front = exp(a * log(x) + b * log(xc) - lbeta);
    if (x < (a + 1.0) / (a + b + 2.0)) return front * betacf(a, b, x) / a;
    return 1.0 - front * betacf(b, a, xc) / b;
}

// survival function of the F distribution: fdtrc(dfn, dfd, f)
__device__ double f_survival(double dfn, double dfd, double f) {
    if (isnan(f) || f < 0.0) return nan("");
    if (isinf(f)) return 0.0;
    double den = dfd + dfn * f;
    return incbeta(0.5 * dfd, 0.5 * dfn, dfd / den, dfn * f / den);
}

struct GroupCounts { double n[16]; };

template <int GMAX>
__global__ void __launch_bounds__(kAnovaThreads)
anova_f_kernel(const float* __restrict__ ea, int64_t Na, const float* __restrict__ eb, int64_t Nb,
               int64_t CL, const int32_t* __restrict__ group, GroupCounts cnt, int G,
               double* __restrict__ Fout, double* __restrict__ Pout) {
    __shared__ unsigned char lab[kLabelChunk];
    const int64_t idx = (int64_t)blockIdx.x * kAnovaThreads + threadIdx.x;
    const bool live = idx < CL;
    const int64_t N = Na + Nb;
    double s[GMAX];
    float mn[GMAX], mx[GMAX];
#pragma unroll
    for (int k = 0; k < GMAX; ++k) { s[k] = 0.0; mn[k] = INFINITY; mx[k] = -INFINITY; }
    double q = 0.0;
    const double shift = live ? (double)(Na > 0 ? ea[idx] : eb[idx]) : 0.0;

    for (int64_t n0 = 0; n0 < N; n0 += kLabelChunk) {
        const int chunk = (int)(N - n0 < kLabelChunk ? N - n0 : kLabelChunk);
        __syncthreads();
        for (int i = threadIdx.x; i < chunk; i += kAnovaThreads) lab[i] = (unsigned char)group[n0 + i];
        __syncthreads();
        if (!live) continue;
#pragma unroll 4
        for (int i = 0; i < chunk; ++i) {
            const int64_t n = n0 + i;
            const float v = n < Na ? __ldg(ea + n * CL + idx) : __ldg(eb + (n - Na) * CL + idx);
            const int g = lab[i];
            const double d = (double)v - shift;
            q = fma(d, d, q);
#pragma unroll
            for (int k = 0; k < GMAX; ++k) {
                const bool hit = (g == k);
                s[k] += hit ? d : 0.0;
                mn[k] = hit ? fminf(mn[k], v) : mn[k];
                mx[k] = hit ? fmaxf(mx[k], v) : mx[k];
            }
        }
    }
    if (!live) return;
    double S = 0.0, ssb = 0.0;
    bool all_const = true;
    float gmn = INFINITY, gmx = -INFINITY;
#pragma unroll
    for (int k = 0; k < GMAX; ++k) {
        if (k < G) {
            S += s[k];
            ssb += s[k] * s[k] / cnt.n[k];
            all_const = all_const && (mn[k] == mx[k]);
            gmn = fminf(gmn, mn[k]); gmx = fmaxf(gmx, mx[k]);
        }
    }
    const double norm = S * S / (double)N;
    const double sstot = q - norm;
    ssb -= norm;
    const double ssw = sstot - ssb;
    const double dfb = (double)(G - 1), dfw = (double)(N - G);
    double F = (ssb / dfb) / (ssw / dfw);
    if (all_const) F = INFINITY;
    if (gmn == gmx) F = nan("");
    Fout[idx] = F;
    Pout[idx] = f_survival(dfb, dfw, F);
}

__global__ void __launch_bounds__(128)
sig_runlength_kernel(const double* __restrict__ p, int C, int64_t L, double thr, int32_t* __restrict__ maxrun) {
    const int lane = threadIdx.x & 31;
    const int ch = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (ch >= C) return;
    const double* row = p + (int64_t)ch * L;
    int best = 0, cur = 0;
    for (int64_t t0 = 0; t0 < L; t0 += 32) {
        const int64_t t = t0 + lane;
        const bool sig = (t < L) && (row[t] < thr);           // NaN compares false
        unsigned m = __ballot_sync(0xffffffffu, sig);
        // every lane walks the same 32-bit mask (uniform, no divergence)
        if (m == 0xffffffffu) { cur += 32; best = cur > best ? cur : best; continue; }
        for (int b = 0; b < 32; ++b) {
            if ((m >> b) & 1u) { ++cur; best = cur > best ? cur : best; }
            else cur = 0;
        }
    }
    if (lane == 0) maxrun[ch] = best;
}

}  // namespace ecog

using namespace ecog;

extern "C" int ecog_anova_f(const float* d_epochs_a, int64_t Na, const float* d_epochs_b, int64_t Nb,
                            int64_t C, int64_t L, const int32_t* d_group, const int64_t* h_group_count,
                            int32_t G, double* d_F, double* d_p, ecog_stream_t stream) {
    if (C <= 0 || L <= 0 || Na < 0 || Nb < 0 || Na + Nb < 2) return fail(ECOG_E_VALUE, "ecog_anova_f: bad shape");
    if (G < 2) return fail(ECOG_E_VALUE, "ecog_anova_f: need at least two groups, got %d", G);
    if (G > 16) return fail(ECOG_E_UNSUPPORTED, "ecog_anova_f: at most 16 groups supported, got %d", G);
    if (Na + Nb <= G) return fail(ECOG_E_VALUE, "ecog_anova_f: need more events than groups");
    GroupCounts cnt;
    int64_t tot = 0;
    for (int k = 0; k < 16; ++k) {
        cnt.n[k] = k < G ? (double)h_group_count[k] : 1.0;
        if (k < G) {
            if (h_group_count[k] <= 0) return fail(ECOG_E_VALUE, "ecog_anova_f: empty group %d", k);
            tot += h_group_count[k];
        }
    }
    if (tot != Na + Nb) return fail(ECOG_E_VALUE, "ecog_anova_f: group counts do not sum to the event count");
    const int64_t CL = C * L;
    unsigned grid = (unsigned)ceil_div(CL, kAnovaThreads);
    cudaStream_t st = (cudaStream_t)stream;
    if (G <= 2) anova_f_kernel<2><<<grid, kAnovaThreads, 0, st>>>(d_epochs_a, Na, d_epochs_b, Nb, CL, d_group, cnt, G, d_F, d_p);
    else if (G <= 4) anova_f_kernel<4><<<grid, kAnovaThreads, 0, st>>>(d_epochs_a, Na, d_epochs_b, Nb, CL, d_group, cnt, G, d_F, d_p);
    else if (G <= 8) anova_f_kernel<8><<<grid, kAnovaThreads, 0, st>>>(d_epochs_a, Na, d_epochs_b, Nb, CL, d_group, cnt, G, d_F, d_p);
    else anova_f_kernel<16><<<grid, kAnovaThreads, 0, st>>>(d_epochs_a, Na, d_epochs_b, Nb, CL, d_group, cnt, G, d_F, d_p);
    return check_launch("anova_f");
}

extern "C" int ecog_sig_runlength(const double* d_p, int64_t C, int64_t L, double threshold,
                                  int32_t* d_maxrun, ecog_stream_t stream) {
    if (C <= 0 || L <= 0) return fail(ECOG_E_VALUE, "ecog_sig_runlength: bad shape");
    sig_runlength_kernel<<<(unsigned)ceil_div(C, 4), 128, 0, (cudaStream_t)stream>>>(d_p, (int)C, L, threshold, d_maxrun);
    return check_launch("sig_runlength");
}
