// K9 per-(channel, timepoint) one-way ANOVA and K10 longest significant run.
//
// anova_f: one thread per (c, t) and event slab; the event loop streams the (N, C, L) epoch
// tensor with stride C*L (coalesced along t), float64 accumulators about the first event's
// value (the F statistic is shift invariant; scipy centres on the grand mean for the
// same reason).  Events are cut into slabs (grid.y) so that enough loads are in flight to
// reach HBM bandwidth when C*L alone is ~1e5 threads; the slab partials are combined in slab
// order by a second kernel (deterministic).  4 B read per epoch element -> HBM-bound.  p = fdtrc(G-1, N-G, F) is
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

// Partial accumulators of one event slab for one (channel, timepoint):
//   sums[GMAX] (float64, about the shift), q = sum d^2 (float64), mn/mx[GMAX] (float32).
// Stored as planes over idx so that the finalise pass reads coalesced.
struct GroupStarts { int64_t at[17]; };     // events of group k are order[at[k] .. at[k+1])

// Events arrive SORTED BY GROUP (d_order: stable argsort of the labels, built by the host), so a
// thread accumulates one group at a time into scalars: load, convert, subtract, add, fma,
// min, max -- no per-group selects.  A slab is a contiguous range of the sorted order.
template <int GMAX>
__global__ void __launch_bounds__(kAnovaThreads)
anova_partial_kernel(const float* __restrict__ ea, int64_t Na, const float* __restrict__ eb, int64_t Nb,
                     int64_t CL, const int32_t* __restrict__ order, GroupStarts gs, int64_t slab,
                     double* __restrict__ psum, float* __restrict__ pmin, float* __restrict__ pmax) {
    __shared__ int32_t ord[kLabelChunk];
    const int64_t idx = (int64_t)blockIdx.x * kAnovaThreads + threadIdx.x;
    const bool live = idx < CL;
    const int64_t N = Na + Nb;
    const int64_t nBeg = (int64_t)blockIdx.y * slab;
    const int64_t nEnd = nBeg + slab < N ? nBeg + slab : N;
    double s[GMAX];
    float mn[GMAX], mx[GMAX];
    double q = 0.0;
    const double shift = live ? (double)(Na > 0 ? ea[idx] : eb[idx]) : 0.0;
    const float* pa = ea + (live ? idx : 0);
    const float* pb = eb ? eb + (live ? idx : 0) - Na * CL : pa;      // event n >= Na lives at eb[(n - Na) * CL]
#pragma unroll
    for (int k = 0; k < GMAX; ++k) {
        double sk = 0.0;
        float lo = INFINITY, hi = -INFINITY;
        const int64_t g0 = gs.at[k] > nBeg ? gs.at[k] : nBeg;
        const int64_t g1 = gs.at[k + 1] < nEnd ? gs.at[k + 1] : nEnd;
        for (int64_t n0 = g0; n0 < g1; n0 += kLabelChunk) {
            const int chunk = (int)(g1 - n0 < kLabelChunk ? g1 - n0 : kLabelChunk);
            __syncthreads();
            for (int i = threadIdx.x; i < chunk; i += kAnovaThreads) ord[i] = order[n0 + i];
            __syncthreads();
            if (!live) continue;
#pragma unroll 8
            for (int i = 0; i < chunk; ++i) {
                const int64_t n = ord[i];
                const float v = __ldg((n < Na ? pa : pb) + n * CL);
                const double d = (double)v - shift;
                sk += d;
                q = fma(d, d, q);
                lo = fminf(lo, v);
                hi = fmaxf(hi, v);
            }
        }
        s[k] = sk; mn[k] = lo; mx[k] = hi;
    }
    if (!live) return;
    const int64_t base = (int64_t)blockIdx.y * (GMAX + 1) * CL;
#pragma unroll
    for (int k = 0; k < GMAX; ++k) psum[base + k * CL + idx] = s[k];
    psum[base + GMAX * CL + idx] = q;
    const int64_t fb = (int64_t)blockIdx.y * GMAX * CL;
#pragma unroll
    for (int k = 0; k < GMAX; ++k) { pmin[fb + k * CL + idx] = mn[k]; pmax[fb + k * CL + idx] = mx[k]; }
}

// combine the slabs in slab order (deterministic) and finish F, p
template <int GMAX>
__global__ void __launch_bounds__(kAnovaThreads)
anova_final_kernel(int64_t CL, int64_t N, int nslab, GroupCounts cnt, int G,
                   const double* __restrict__ psum, const float* __restrict__ pmin, const float* __restrict__ pmax,
                   double* __restrict__ Fout, double* __restrict__ Pout) {
    const int64_t idx = (int64_t)blockIdx.x * kAnovaThreads + threadIdx.x;
    if (idx >= CL) return;
    double s[GMAX];
    float mn[GMAX], mx[GMAX];
#pragma unroll
    for (int k = 0; k < GMAX; ++k) { s[k] = 0.0; mn[k] = INFINITY; mx[k] = -INFINITY; }
    double q = 0.0;
    for (int b = 0; b < nslab; ++b) {
        const int64_t base = (int64_t)b * (GMAX + 1) * CL, fb = (int64_t)b * GMAX * CL;
#pragma unroll
        for (int k = 0; k < GMAX; ++k) {
            s[k] += psum[base + k * CL + idx];
            mn[k] = fminf(mn[k], pmin[fb + k * CL + idx]);
            mx[k] = fmaxf(mx[k], pmax[fb + k * CL + idx]);
        }
        q += psum[base + GMAX * CL + idx];
    }
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

// ---- more than 16 groups (e.g. a syllable target with many classes; scipy's f_oneway has no limit):
// same algorithm with the group loop at run time -- each group's partials go straight to the
// workspace instead of a register array.  Group starts ride in the kernel parameters.
constexpr int kMaxGroups = 256;
struct GroupStartsBig { int64_t at[kMaxGroups + 1]; };

__global__ void __launch_bounds__(kAnovaThreads)
anova_partial_rt_kernel(const float* __restrict__ ea, int64_t Na, const float* __restrict__ eb, int64_t Nb,
                        int64_t CL, const int32_t* __restrict__ order, GroupStartsBig gs, int G, int64_t slab,
                        double* __restrict__ psum, float* __restrict__ pmin, float* __restrict__ pmax) {
    __shared__ int32_t ord[kLabelChunk];
    const int64_t idx = (int64_t)blockIdx.x * kAnovaThreads + threadIdx.x;
    const bool live = idx < CL;
    const int64_t N = Na + Nb;
    const int64_t nBeg = (int64_t)blockIdx.y * slab;
    const int64_t nEnd = nBeg + slab < N ? nBeg + slab : N;
    double q = 0.0;
    const double shift = live ? (double)(Na > 0 ? ea[idx] : eb[idx]) : 0.0;
    const float* pa = ea + (live ? idx : 0);
    const float* pb = eb ? eb + (live ? idx : 0) - Na * CL : pa;
    const int64_t base = (int64_t)blockIdx.y * (G + 1) * CL;
    const int64_t fb = (int64_t)blockIdx.y * G * CL;
    for (int k = 0; k < G; ++k) {
        double sk = 0.0;
        float lo = INFINITY, hi = -INFINITY;
        const int64_t g0 = gs.at[k] > nBeg ? gs.at[k] : nBeg;
        const int64_t g1 = gs.at[k + 1] < nEnd ? gs.at[k + 1] : nEnd;
        for (int64_t n0 = g0; n0 < g1; n0 += kLabelChunk) {
            const int chunk = (int)(g1 - n0 < kLabelChunk ? g1 - n0 : kLabelChunk);
            __syncthreads();
            for (int i = threadIdx.x; i < chunk; i += kAnovaThreads) ord[i] = order[n0 + i];
            __syncthreads();
            if (!live) continue;
#pragma unroll 8
            for (int i = 0; i < chunk; ++i) {
                const int64_t n = ord[i];
                const float v = __ldg((n < Na ? pa : pb) + n * CL);
                const double d = (double)v - shift;
                sk += d;
                q = fma(d, d, q);
                lo = fminf(lo, v);
                hi = fmaxf(hi, v);
            }
        }
        if (live) { psum[base + k * CL + idx] = sk; pmin[fb + k * CL + idx] = lo; pmax[fb + k * CL + idx] = hi; }
    }
    if (live) psum[base + (int64_t)G * CL + idx] = q;
}

__global__ void __launch_bounds__(kAnovaThreads)
anova_final_rt_kernel(int64_t CL, int64_t N, int nslab, GroupStartsBig gs, int G,
                      const double* __restrict__ psum, const float* __restrict__ pmin, const float* __restrict__ pmax,
                      double* __restrict__ Fout, double* __restrict__ Pout) {
    const int64_t idx = (int64_t)blockIdx.x * kAnovaThreads + threadIdx.x;
    if (idx >= CL) return;
    double S = 0.0, ssb = 0.0, q = 0.0;
    bool all_const = true;
    float gmn = INFINITY, gmx = -INFINITY;
    for (int k = 0; k < G; ++k) {
        double sk = 0.0;
        float mn = INFINITY, mx = -INFINITY;
        for (int b = 0; b < nslab; ++b) {           // slab order: deterministic
            sk += psum[(int64_t)b * (G + 1) * CL + k * CL + idx];
            mn = fminf(mn, pmin[(int64_t)b * G * CL + k * CL + idx]);
            mx = fmaxf(mx, pmax[(int64_t)b * G * CL + k * CL + idx]);
        }
        S += sk;
        ssb += sk * sk / (double)(gs.at[k + 1] - gs.at[k]);
        all_const = all_const && (mn == mx);
        gmn = fminf(gmn, mn); gmx = fmaxf(gmx, mx);
    }
    for (int b = 0; b < nslab; ++b) q += psum[(int64_t)b * (G + 1) * CL + (int64_t)G * CL + idx];
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

static int gmax_of(int G) { return G <= 2 ? 2 : (G <= 4 ? 4 : (G <= 8 ? 8 : (G <= 16 ? 16 : G))); }
static int anova_slabs(int64_t CL, int64_t N, int G) {
    // enough (c,t)-threads x slabs to fill the machine with loads in flight; slabs of >= 256 events
    int64_t want = ceil_div((int64_t)kNumSMs * 2048 * 2, CL);
    int64_t cap = N / 256 > 1 ? N / 256 : 1;
    int64_t sl = want < cap ? want : cap;
    if (G > 16) {   // run-time group path: keep the partials below ~1 GiB
        const int64_t per = CL * ((int64_t)(G + 1) * 8 + 2 * (int64_t)G * 4);
        const int64_t fit = ((int64_t)1 << 30) / (per > 0 ? per : 1);
        if (sl > fit) sl = fit;
    }
    return (int)(sl < 1 ? 1 : (sl > 64 ? 64 : sl));
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

extern "C" size_t ecog_anova_workspace(int64_t C, int64_t L, int64_t N, int32_t G) {
    const int64_t CL = C * L;
    const int gm = gmax_of(G), ns = anova_slabs(CL, N, G);
    return (size_t)ns * CL * ((gm + 1) * sizeof(double) + 2 * gm * sizeof(float)) + 256;
}

extern "C" int ecog_anova_f(const float* d_epochs_a, int64_t Na, const float* d_epochs_b, int64_t Nb,
                            int64_t C, int64_t L, const int32_t* d_order, const int64_t* h_group_count,
                            int32_t G, double* d_F, double* d_p, void* d_workspace, size_t workspace_bytes,
                            ecog_stream_t stream) {
    if (C <= 0 || L <= 0 || Na < 0 || Nb < 0 || Na + Nb < 2) return fail(ECOG_E_VALUE, "ecog_anova_f: bad shape");
    if (G < 2) return fail(ECOG_E_VALUE, "ecog_anova_f: need at least two groups, got %d", G);
    if (G > kMaxGroups) return fail(ECOG_E_UNSUPPORTED, "ecog_anova_f: at most %d groups supported, got %d", kMaxGroups, G);
    if (Na + Nb <= G) return fail(ECOG_E_VALUE, "ecog_anova_f: need more events than groups");
    if (G > 16) {
        GroupStartsBig gb;
        int64_t t2 = 0;
        for (int k = 0; k <= kMaxGroups; ++k) {
            gb.at[k] = t2;
            if (k < G) {
                if (h_group_count[k] <= 0) return fail(ECOG_E_VALUE, "ecog_anova_f: empty group %d", k);
                t2 += h_group_count[k];
            }
        }
        const int64_t N2 = Na + Nb;
        if (t2 != N2) return fail(ECOG_E_VALUE, "ecog_anova_f: group counts do not sum to the event count");
        if (workspace_bytes < ecog_anova_workspace(C, L, N2, G))
            return fail(ECOG_E_WORKSPACE, "ecog_anova_f: workspace %zu < %zu", workspace_bytes, ecog_anova_workspace(C, L, N2, G));
        const int64_t CL2 = C * L;
        const int ns2 = anova_slabs(CL2, N2, G);
        const int64_t slab2 = ceil_div(N2, ns2);
        double* ps = (double*)d_workspace;
        float* pmn = (float*)(ps + (size_t)ns2 * (G + 1) * CL2);
        float* pmx = pmn + (size_t)ns2 * G * CL2;
        dim3 g2((unsigned)ceil_div(CL2, kAnovaThreads), (unsigned)ns2);
        cudaStream_t s2 = (cudaStream_t)stream;
        anova_partial_rt_kernel<<<g2, kAnovaThreads, 0, s2>>>(d_epochs_a, Na, d_epochs_b, Nb, CL2, d_order, gb, G, slab2,
                                                             ps, pmn, pmx);
        ECOG_TRY(check_launch("anova_partial_rt"));
        anova_final_rt_kernel<<<(unsigned)ceil_div(CL2, kAnovaThreads), kAnovaThreads, 0, s2>>>(CL2, N2, ns2, gb, G, ps, pmn,
                                                                                              pmx, d_F, d_p);
        return check_launch("anova_final_rt");
    }
    GroupCounts cnt;
    GroupStarts gs;
    int64_t tot = 0;
    for (int k = 0; k < 16; ++k) {
        cnt.n[k] = k < G ? (double)h_group_count[k] : 1.0;
        gs.at[k] = tot;
        if (k < G) {
            if (h_group_count[k] <= 0) return fail(ECOG_E_VALUE, "ecog_anova_f: empty group %d", k);
            tot += h_group_count[k];
        }
    }
    gs.at[16] = tot;
    const int64_t N = Na + Nb;
    if (tot != N) return fail(ECOG_E_VALUE, "ecog_anova_f: group counts do not sum to the event count");
    if (workspace_bytes < ecog_anova_workspace(C, L, N, G))
        return fail(ECOG_E_WORKSPACE, "ecog_anova_f: workspace %zu < %zu", workspace_bytes, ecog_anova_workspace(C, L, N, G));
    const int64_t CL = C * L;
    const int gm = gmax_of(G), ns = anova_slabs(CL, N, G);
    const int64_t slab = ceil_div(N, ns);
    double* psum = (double*)d_workspace;
    float* pmin = (float*)(psum + (size_t)ns * (gm + 1) * CL);
    float* pmax = pmin + (size_t)ns * gm * CL;
    dim3 grid((unsigned)ceil_div(CL, kAnovaThreads), (unsigned)ns);
    const unsigned fgrid = (unsigned)ceil_div(CL, kAnovaThreads);
    cudaStream_t st = (cudaStream_t)stream;
#define ECOG_ANOVA(GM)                                                                                              \
    do {                                                                                                            \
        anova_partial_kernel<GM><<<grid, kAnovaThreads, 0, st>>>(d_epochs_a, Na, d_epochs_b, Nb, CL, d_order, gs, slab, \
                                                                 psum, pmin, pmax);                                 \
        ECOG_TRY(check_launch("anova_partial"));                                                                    \
        anova_final_kernel<GM><<<fgrid, kAnovaThreads, 0, st>>>(CL, N, ns, cnt, G, psum, pmin, pmax, d_F, d_p);     \
    } while (0)
    if (gm == 2) ECOG_ANOVA(2);
    else if (gm == 4) ECOG_ANOVA(4);
    else if (gm == 8) ECOG_ANOVA(8);
    else ECOG_ANOVA(16);
#undef ECOG_ANOVA
    return check_launch("anova_final");
}

extern "C" int ecog_sig_runlength(const double* d_p, int64_t C, int64_t L, double threshold,
                                  int32_t* d_maxrun, ecog_stream_t stream) {
    if (C <= 0 || L <= 0) return fail(ECOG_E_VALUE, "ecog_sig_runlength: bad shape");
    sig_runlength_kernel<<<(unsigned)ceil_div(C, 4), 128, 0, (cudaStream_t)stream>>>(d_p, (int)C, L, threshold, d_maxrun);
    return check_launch("sig_runlength");
}
