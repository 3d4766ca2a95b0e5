// Shared pieces of the IIR biquad-cascade kernels (K3): constants, the cascade step and the
// warm-up sweep kernel, used by sosfilt.cu (single cascades, scan path, entry points) and
// sosfilt_pair.cu (two cascades fused into one sweep pair).
#pragma once
#include "common.cuh"

namespace ecog {

constexpr int kSosThreads = 512;
constexpr int kSub = 16;              // samples per stage per chunk
constexpr int kPitch = kSub + 4;      // shared row pitch in floats (conflict-free LDS.128)
constexpr int kRing = 2;
constexpr int kWarmRing = 5;      // warm-up kernel: tile slots (prefetch distance kWarmRing - 1), one barrier per stage
constexpr int kWSub = 16;         // warm-up kernel: samples per stage per chunk (64 B per chunk and stage; 128 B measured no faster)
constexpr int kWPitch = kWSub;   // no padding: the 16-byte pieces of a row are XOR-swizzled (conflict-free LDS.128)

struct SosCoef {
    double c[ECOG_MAX_SECTIONS][5];   // b0 b1 b2 a1 a2
    double zi[ECOG_MAX_SECTIONS][2];
};
struct SosMatrix { double m[2 * ECOG_MAX_SECTIONS][2 * ECOG_MAX_SECTIONS]; };

// Numerator forms (NUM):
//   0  general b0 b1 b2                                   5 FP64 ops per section and sample
//   1  b1 == 0                                            4
//   2  unit form  g * (1 + beta1 z^-1 + z^-2) per section: the gain g of the cascade is applied
//      once to the input, sections are monic with b2 == +1 (Butterworth band-stop / low-pass /
//      high-pass: zeros on the unit circle)               4, and one multiply per sample
//   5  unit form with b1 == 0 and b2 == -1, the (1 - z^-2) sections of a Butterworth band-pass
//                                                         3, and one multiply per sample
//   8  monic general numerator 1 + b1 z^-1 + b2 z^-2 in DIRECT FORM II (the exact factors of a rounded
//      Butterworth numerator whose zeros are NOT on the unit circle; gain on the input):
//        w = u - a1 w1 - a2 w2;  y = w + b1 w1 + b2 w2 = u + (b1 - a1) w1 + (b2 - a2) w2      4, and one multiply per sample
//      (the transposed form needs 5: its b2 u product has no partner).  In float64 the direct form is as
//      accurate as the transposed one for these sections (both 1.5e-7 from the long-double evaluation of
//      the 58-62 Hz notch at 3 kHz); its states (w1, w2) are related to the transposed ones by
//      s0 = (b1 - a1) w1 + (b2 - a2) w2,  s1 = (b2 - a2) w1 + (b2 a1 - a2 b1) w2  (df2_state below).
// In the unit forms c[j][1] holds beta1 = b1 / b0; the states are those of the general form.
template <int J0, int J1, int NUM, int NSEC>
__device__ __forceinline__ double sos_range(double u, const double (&c)[NSEC][5], double (&s)[NSEC][2]) {
    constexpr bool B1Z = (NUM & 1) != 0;
    constexpr int UNIT = NUM >> 1;              // 0 general, 1: b0 = 1, b2 = +1, 2: b0 = 1, b2 = -1
    if (NUM == 8) {
#pragma unroll
        for (int j = J0; j < J1; ++j) {
            // the newest state w1 enters LAST (one dependent DFMA, ~15 cycles, per sample on the recurrence), and
            // the output is formed beside it, not after it:  y = w + b1 w1 + b2 w2 = u + (b1 - a1) w1 + (b2 - a2) w2
            // (prepare_form stores the differences in c[j][1], c[j][2])
            const double w = fma(-c[j][3], s[j][0], fma(-c[j][4], s[j][1], u));
            u = fma(c[j][1], s[j][0], fma(c[j][2], s[j][1], u));
            s[j][1] = s[j][0];
            s[j][0] = w;
        }
        return u;
    }
#pragma unroll
    for (int j = J0; j < J1; ++j) {
        const double y = UNIT ? u + s[j][0] : fma(c[j][0], u, s[j][0]);
        s[j][0] = B1Z ? fma(-c[j][3], y, s[j][1]) : fma(-c[j][3], y, fma(c[j][1], u, s[j][1]));
        s[j][1] = UNIT == 1 ? fma(-c[j][4], y, u) : UNIT == 2 ? fma(-c[j][4], y, -u) : fma(-c[j][4], y, c[j][2] * u);
        u = y;
    }
    return u;
}

// NUMB < 0: one numerator form for the whole cascade.  NUMB >= 0: a PAIR of 4-section cascades run
// as one (NSEC == 8): sections 0-3 in form NUM, sections 4-7 in form NUMB (both unit forms, the
// product of the two gains rides on the input); `full` = false runs the first cascade only (early
// warm-up: the second cascade forgets faster and starts later from a zero state).
template <int NSEC, int NUM = 0, int NUMB = -1>
__device__ __forceinline__ double sos_step(double u, const double (&c)[NSEC][5], double (&s)[NSEC][2], bool full = true) {
    if (NUMB < 0) return sos_range<0, NSEC, NUM, NSEC>(u, c, s);
    u = sos_range<0, NSEC / 2, NUM, NSEC>(u, c, s);
    if (full) u = sos_range<NSEC / 2, NSEC, (NUMB < 0 ? 0 : NUMB), NSEC>(u, c, s);
    return u;
}

// N consecutive samples through sections [0, NS) in WAVEFRONT order: on diagonal d section j handles sample
// d - j, so the NS section steps of one diagonal are independent of each other (different sections'
// states, different samples) and sit next to each other in program order.  Same operations as N calls of
// sos_step, another schedule: a thread then has ~NS independent FP64 chains in flight instead of one, which
// is what two warps per scheduler need to keep the FP64 pipe busy (ncu: the sample-by-sample order stalls
// on `wait`, the fixed-latency dependency of consecutive DFMAs).  GAIN: the input is multiplied by `gain`.
template <int NSEC, int NS, int NUM, int NUMB, int N, bool GAIN, bool STORE>
__device__ __forceinline__ void sos_block(const float (&x)[N], float (&y)[N], double gain,
                                          const double (&c)[NSEC][5], double (&s)[NSEC][2]) {
    double pipe[9];
#pragma unroll
    for (int d = 0; d < N + NS - 1; ++d) {
#pragma unroll
        for (int j = NS - 1; j >= 0; --j) {
            const int n = d - j;
            if (n >= 0 && n < N) {
                const double u = j == 0 ? (GAIN ? gain * (double)x[n] : (double)x[n]) : pipe[j];
                if (NUMB < 0 || j < NSEC / 2) {
                    switch (j) {     // compile-time j after unrolling
                        case 0: pipe[1] = sos_range<0, 1, NUM, NSEC>(u, c, s); break;
                        case 1: pipe[2] = sos_range<(NSEC > 1 ? 1 : 0), (NSEC > 1 ? 2 : 1), NUM, NSEC>(u, c, s); break;
                        case 2: pipe[3] = sos_range<(NSEC > 2 ? 2 : 0), (NSEC > 2 ? 3 : 1), NUM, NSEC>(u, c, s); break;
                        case 3: pipe[4] = sos_range<(NSEC > 3 ? 3 : 0), (NSEC > 3 ? 4 : 1), NUM, NSEC>(u, c, s); break;
                        case 4: pipe[5] = sos_range<(NSEC > 4 ? 4 : 0), (NSEC > 4 ? 5 : 1), NUM, NSEC>(u, c, s); break;
                        case 5: pipe[6] = sos_range<(NSEC > 5 ? 5 : 0), (NSEC > 5 ? 6 : 1), NUM, NSEC>(u, c, s); break;
                        case 6: pipe[7] = sos_range<(NSEC > 6 ? 6 : 0), (NSEC > 6 ? 7 : 1), NUM, NSEC>(u, c, s); break;
                        default: pipe[8] = sos_range<(NSEC > 7 ? 7 : 0), (NSEC > 7 ? 8 : 1), NUM, NSEC>(u, c, s); break;
                    }
                } else {
                    constexpr int NB = NUMB < 0 ? 0 : NUMB;
                    switch (j) {
                        case 4: pipe[5] = sos_range<(NSEC > 4 ? 4 : 0), (NSEC > 4 ? 5 : 1), NB, NSEC>(u, c, s); break;
                        case 5: pipe[6] = sos_range<(NSEC > 5 ? 5 : 0), (NSEC > 5 ? 6 : 1), NB, NSEC>(u, c, s); break;
                        case 6: pipe[7] = sos_range<(NSEC > 6 ? 6 : 0), (NSEC > 6 ? 7 : 1), NB, NSEC>(u, c, s); break;
                        default: pipe[8] = sos_range<(NSEC > 7 ? 7 : 0), (NSEC > 7 ? 8 : 1), NB, NSEC>(u, c, s); break;
                    }
                }
            }
        }
        if (STORE && d >= NS - 1) y[d - (NS - 1)] = (float)pipe[NS];
    }
}

// ------------------------------------------------------------------ float32 band-pass half of a cascade pair
// The second half of a pair (four (1 - z^-2) sections of a Butterworth band-pass, poles well inside the unit
// circle) in FLOAT32, in delta form: with a1 = e1 - 2, a2 = 1 - e2 the direct-form-II recursion
//     w[n] = v - a1 w[n-1] - a2 w[n-2],   y = w[n] - w[n-2]
// is carried in the states w1 = w[n-1] and d = w[n-1] - w[n-2]:
//     dn = d + (v + (e2 - e1) w1 - e2 d),   y = dn + d,   w1 += dn,   d = dn          (2 FFMA + 3 FADD)
// The small coefficients c1 = e2 - e1 = -(1 + a1 + a2) and e2 = 1 - a2 are rounded to float32 with an absolute
// error ~1e-9 instead of the 6e-8 of a1 ~ -1.9, and the recursion adds small increments to the state instead
// of cancelling large products: measured against the all-float64 pair 3.5e-7 (2 kHz) / 4.5e-7 (3 kHz) of the
// row maximum for the 70-150 Hz band, where plain float32 direct forms give 2.5e-6 ... 5e-6.
// It takes 12 of the pair's 29 FP64 operations per sample off the FP64 pipe that bounds the sweeps
// (csrc/sosfilt_pairws.cu runs the recursion, two sections per packed float32 pair).
struct Bp32Coef { float c1[4], e2[4]; };

// WRITE=false: tail pass (zero state, last `tail` samples, end state -> slot k+1)
// WRITE=true : main pass (state from slot k, float32 output)
// ------------------------------------------------------------------ warm-up path
// Single kernel per sweep, no scan: when the cascade forgets a zero-state start within
// `tail` samples (max |A^tail| < 1e-10, decided by the host) a chunk's true start state is
// reproduced by running the recurrence from a ZERO state over the `tail` samples that precede
// the chunk.  One thread = one chunk: stages -tail/16 .. -1 are the warm-up (loads only),
// stages 0 .. L/16-1 filter the chunk and are written.  Chunks whose warm-up would cross the
// row edge are exact instead: the filtfilt start-up state (zi * ext[0] pushed through the odd
// extension pad) is injected at the stage where the row starts.  Redundant work is tail/L,
// against a full extra pass for the scan path, and the chunk length is chosen for a few
// warps per scheduler only (the FP64 pipe saturates early), which makes L long.
// Thread q -> (row q / nChunks, chunk q % nChunks): a CTA walks neighbouring chunks of one row
// (few DRAM pages / TLB entries live per CTA).
// The backward sweep cannot run in place (its warm-up reads the forward result of the
// neighbouring chunk), so the forward result lives in the workspace.
template <int NSEC, bool REV, bool VEC, int NT, int NUM, int NUMB = -1>
__global__ void __launch_bounds__(NT, (NSEC >= 8 ? 1 : 512 / NT))     // 8 sections: 64 coefficient + 32 state registers
sos_warm_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t C, int64_t T,
                int64_t ldx, int64_t ldy, int L, int tail, int nChunks, int padlen, int zero_phase,
                SosCoef coef, double* __restrict__ padbuf, double gain, int tail_b) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* tiles = reinterpret_cast<float*>(smem_raw);                                   // [kWarmRing][NT][kWPitch]
    int64_t* soff = reinterpret_cast<int64_t*>(tiles + (size_t)kWarmRing * NT * kWPitch); // [NT] x offset of the chunk edge
    int64_t* doff = soff + NT;                                                           // [NT] y offset of the chunk edge
    int2* lohi = reinterpret_cast<int2*>(doff + NT);                                     // [NT] valid logical offsets
    constexpr int PE = 4;                       // samples per piece (one 16 B copy when VEC)
    constexpr int PP = kWSub / PE;              // pieces per tile row
    constexpr int RPB = 32 / kWSub;             // tile rows per 128 bytes of shared memory (swizzle period)
    constexpr int NJ = PP;                      // pieces each thread moves per stage

    const int tid = threadIdx.x;
    const int64_t q = (int64_t)blockIdx.x * NT + tid;
    const bool valid = q < C * nChunks;
    const int64_t row = valid ? q / nChunks : 0;
    const int k = valid ? (int)(q - row * nChunks) : 0;
    int64_t edge; int ulo, uhi;
    {
        int64_t a, b;
        if (!REV) { a = (int64_t)k * L; b = a + L < T ? a + L : T; }
        else      { b = T - (int64_t)k * L; a = b - L > 0 ? b - L : 0; }
        const int64_t bef = REV ? T - b : a;            // samples between the row edge and the chunk, sweep order
        edge = REV ? b : a;
        ulo = valid ? (bef < tail ? -(int)bef : -tail) : 0;
        uhi = valid ? (int)(b - a) : 0;
    }
    soff[tid] = row * ldx + edge;
    doff[tid] = row * ldy + edge;
    lohi[tid] = make_int2(ulo, uhi);
    __syncthreads();
    const int pc = tid % PP;                    // this thread moves piece pc of tile rows tid / PP + j * (NT / PP)

    double c[NSEC][5], s[NSEC][2];
    // unit forms take the cascade gain on the input
#define IN(v) ((NUM >> 1) ? gain * (v) : (v))
#pragma unroll
    for (int j = 0; j < NSEC; ++j) {
#pragma unroll
        for (int i = 0; i < 5; ++i) c[j][i] = coef.c[j][i];
        s[j][0] = 0.0; s[j][1] = 0.0;
    }
    // chunks that see the row edge get the exact start-up at the stage where the row starts
    const int64_t before = REV ? T - edge : edge;
    const int s_inject = valid && before <= tail ? -(int)(before / kWSub) : (1 << 30);
    // cascade pair: the second cascade joins the warm-up tail_b samples before the chunk (or at the
    // exact start-up of a chunk that sees the row edge)
    const int s_full = NUMB < 0 ? -(1 << 30) : (s_inject < -(tail_b / kWSub) ? s_inject : -(tail_b / kWSub));

    const int nStages = L / kWSub;
    const int first = -(tail / kWSub);

    // logical offset of the first element of piece pc in stage st, and its distance from the chunk edge
    auto issue = [&](int stage) {
        if (stage < nStages) {
            float* tile = tiles + (size_t)((stage - first) % kWarmRing) * NT * kWPitch;
            const int u0 = !REV ? stage * kWSub + PE * pc : stage * kWSub + (kWSub - PE) - PE * pc;
            const int off = !REV ? u0 : -u0 - PE;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int r = tid / PP + j * (NT / PP);
                const int2 lh = lohi[r];
                const float* g = x + soff[r];
                float* d = tile + r * kWPitch + PE * (pc ^ ((r / RPB) & (PP - 1)));
                if (VEC) {
                    const bool ok = u0 >= lh.x && u0 + PE <= lh.y;
                    cp_async16_zfill(d, ok ? g + off : g, ok);
                } else {
#pragma unroll
                    for (int e = 0; e < PE; ++e) {
                        const int u = !REV ? u0 + e : u0 + PE - 1 - e;      // element e in ascending address order
                        const bool ok = u >= lh.x && u < lh.y;
                        cp_async4_zfill(d + e, ok ? g + off + e : g, ok);
                    }
                }
            }
        }
        cp_async_commit();
    };

    // kWarmRing tile slots, ONE barrier per stage: the barrier that publishes the results of stage st
    // also publishes the tiles of stage st + 1 (every thread has waited for its own copies); the
    // slot refilled after it (stage st + kWarmRing - 1) was last read by the store of stage st - 1.
    // Four stages of loads stay in flight per CTA, so neither the copy wait nor the barrier sees
    // DRAM latency.
#pragma unroll
    for (int i = 0; i < kWarmRing - 1; ++i) issue(first + i);
    cp_async_wait<kWarmRing - 2>();
    __syncthreads();
    for (int st = first; st < nStages; ++st) {
        if (st == s_inject && zero_phase) {
            // filtfilt start-up: zi * ext[0], then the odd-extension pad (zero state for the causal filter)
            if (!REV) {
                const float* xr = x + row * ldx;
                const float x0 = xr[0];
                const float e0 = 2.0f * x0 - xr[padlen];
#pragma unroll
                for (int j = 0; j < NSEC; ++j) { s[j][0] = coef.zi[j][0] * (double)e0; s[j][1] = coef.zi[j][1] * (double)e0; }
                for (int i = 0; i < padlen; ++i) {
                    const float e = 2.0f * x0 - xr[padlen - i];
                    (void)sos_step<NSEC, NUM, NUMB>(IN((double)e), c, s);
                }
            } else {
                const double* pb = padbuf + row * padlen;
                const double y0 = pb[padlen - 1];
#pragma unroll
                for (int j = 0; j < NSEC; ++j) { s[j][0] = coef.zi[j][0] * y0; s[j][1] = coef.zi[j][1] * y0; }
                for (int i = padlen - 1; i >= 0; --i) (void)sos_step<NSEC, NUM, NUMB>(IN(pb[i]), c, s);
            }
        }
        float* tile = tiles + (size_t)((st - first) % kWarmRing) * NT * kWPitch;
        float* mine = tile + tid * kWPitch;
        const int swz = (tid / RPB) & (PP - 1);
        float4 xin[kWSub / 4];
#pragma unroll
        for (int v = 0; v < kWSub / 4; ++v) {
            if (!REV) {
                xin[v] = *reinterpret_cast<const float4*>(mine + 4 * (v ^ swz));
            } else {
                float4 t4 = *reinterpret_cast<const float4*>(mine + 4 * ((PP - 1 - v) ^ swz));
                xin[v] = make_float4(t4.w, t4.z, t4.y, t4.x);
            }
        }
        const bool write = st >= 0;
        const int sbase = st * kWSub;
        if (NUMB >= 0 && st < s_full) {                      // early warm-up of a pair: first cascade only, nothing stored
            if (sbase >= ulo && sbase + kWSub <= uhi) {
                float xs[kWSub], ys[kWSub];
#pragma unroll
                for (int v = 0; v < kWSub / 4; ++v) { xs[4 * v] = xin[v].x; xs[4 * v + 1] = xin[v].y; xs[4 * v + 2] = xin[v].z; xs[4 * v + 3] = xin[v].w; }
                sos_block<NSEC, (NSEC / 2 > 0 ? NSEC / 2 : 1), NUM, NUMB, kWSub, (NUM >> 1) != 0, false>(xs, ys, gain, c, s);
            }
        } else if (sbase >= ulo && sbase + kWSub <= uhi) {          // whole stage inside the row: the common case
            float xs[kWSub], ys[kWSub];
#pragma unroll
            for (int v = 0; v < kWSub / 4; ++v) { xs[4 * v] = xin[v].x; xs[4 * v + 1] = xin[v].y; xs[4 * v + 2] = xin[v].z; xs[4 * v + 3] = xin[v].w; }
            sos_block<NSEC, NSEC, NUM, NUMB, kWSub, (NUM >> 1) != 0, true>(xs, ys, gain, c, s);
            if (write) {
#pragma unroll
                for (int v = 0; v < kWSub / 4; ++v) {
                    if (!REV) *reinterpret_cast<float4*>(mine + 4 * (v ^ swz)) = make_float4(ys[4 * v], ys[4 * v + 1], ys[4 * v + 2], ys[4 * v + 3]);
                    else *reinterpret_cast<float4*>(mine + 4 * ((PP - 1 - v) ^ swz)) = make_float4(ys[4 * v + 3], ys[4 * v + 2], ys[4 * v + 1], ys[4 * v]);
                }
            }
        } else {                                            // outside the row / ragged chunk end: state frozen
#pragma unroll
            for (int v = 0; v < kWSub / 4; ++v) {
                const float xv[4] = {xin[v].x, xin[v].y, xin[v].z, xin[v].w};
                float yv[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int u = sbase + 4 * v + e;
                    if (u >= ulo && u < uhi) yv[e] = (float)sos_step<NSEC, NUM, NUMB>(IN((double)xv[e]), c, s);
                }
                if (write) {
                    if (!REV) *reinterpret_cast<float4*>(mine + 4 * (v ^ swz)) = make_float4(yv[0], yv[1], yv[2], yv[3]);
                    else *reinterpret_cast<float4*>(mine + 4 * ((PP - 1 - v) ^ swz)) = make_float4(yv[3], yv[2], yv[1], yv[0]);
                }
            }
        }
        cp_async_wait<kWarmRing - 3>();        // this thread's copies of stage st + 1 have landed
        __syncthreads();
        issue(st + kWarmRing - 1);
        if (write) {
            const int u0 = !REV ? sbase + PE * pc : sbase + (kWSub - PE) - PE * pc;
            const int off = !REV ? u0 : -u0 - PE;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
                const int r = tid / PP + j * (NT / PP);
                const int hi = lohi[r].y;
                float* g = y + doff[r] + off;
                const float4 v4 = *reinterpret_cast<const float4*>(tile + r * kWPitch + PE * (pc ^ ((r / RPB) & (PP - 1))));
                if (VEC) {
                    if (u0 + PE <= hi) *reinterpret_cast<float4*>(g) = v4;
                } else {
                    const float ve[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
                    for (int e = 0; e < PE; ++e) {
                        const int u = !REV ? u0 + e : u0 + PE - 1 - e;
                        if (u < hi) g[e] = ve[e];
                    }
                }
            }
        }
    }
    cp_async_wait<0>();

    // forward sweep: the thread that owns a row's last chunk runs on through the right odd-extension
    // pad (float32 like scipy's odd_ext) and keeps the filtered pad in float64 for the backward start-up
    if (!REV && zero_phase && valid && k == nChunks - 1) {
        const float* xr = x + row * ldx;
        const float xe = xr[T - 1];
        double* pb = padbuf + row * padlen;
        for (int i = 0; i < padlen; ++i) {
            const float e = 2.0f * xe - xr[T - 2 - i];
            pb[i] = sos_step<NSEC, NUM, NUMB>(IN((double)e), c, s);
        }
    }
}
#undef IN


// numerator form of sections [j0, j1): 2 = unit (1 + beta z^-1 + z^-2), 5 = unit (1 - z^-2), 8 = monic
// general (direct form II), 0 = general.  `lead` = the section allowed to carry the gain in b0.
inline int unit_form(const SosCoef& coef, int j0, int j1, int lead) {
    bool b1z = true, unit_p = true, unit_m = true, monic = true;
    for (int j = j0; j < j1; ++j) {
        const double b0 = coef.c[j][0], b2 = coef.c[j][2];
        b1z = b1z && coef.c[j][1] == 0.0;
        const bool b0ok = b0 != 0.0 && (j == lead || b0 == 1.0);
        monic = monic && b0ok;
        unit_p = unit_p && b0ok && b2 == b0;
        unit_m = unit_m && b0ok && b2 == -b0;
    }
    return unit_p ? 2 : (unit_m && b1z ? 5 : (monic ? 8 : 0));
}

// Prepare sections [j0, j1) for their kernel form: forms 2 / 5 / 8 take the gain of section `lead` out
// (returned; the kernel multiplies the input by the product of the gains), form 8 also moves the
// start-up states zi from transposed (s0, s1) to direct-form (w1, w2) coordinates.
inline double prepare_form(SosCoef& coef, int j0, int j1, int lead, int form) {
    double gain = 1.0;
    if (form == 0) return gain;
    if (lead >= j0 && lead < j1) {
        gain = coef.c[lead][0];
        coef.c[lead][1] /= gain;
        if (form == 8) coef.c[lead][2] /= gain;
        // (the unit forms never read c[.][0] and c[.][2])
    }
    if (form == 8) {
        for (int j = j0; j < j1; ++j) {
            const double b1 = coef.c[j][1], b2 = coef.c[j][2], a1 = coef.c[j][3], a2 = coef.c[j][4];
            const double m00 = b1 - a1, m01 = b2 - a2, m10 = b2 - a2, m11 = b2 * a1 - a2 * b1;
            const double det = m00 * m11 - m01 * m10;
            const double s0 = coef.zi[j][0], s1 = coef.zi[j][1];
            if (det != 0.0) {
                coef.zi[j][0] = (m11 * s0 - m01 * s1) / det;
                coef.zi[j][1] = (m00 * s1 - m10 * s0) / det;
            }
            // the kernel forms the output as u + (b1 - a1) w1 + (b2 - a2) w2 (sos_range, NUM == 8)
            coef.c[j][1] = b1 - a1;
            coef.c[j][2] = b2 - a2;
        }
    }
    return gain;
}

// defined in sosfilt_tma.cu
int run_sos_warm_tma(const float* x, float* y, int64_t C, int64_t T, const ecog_sos_plan& p, const SosCoef& coef_in,
                     float* tmp, double* padbuf, cudaStream_t st);

// defined in sosfilt_pairws.cu
int run_sos_pair_ws(const float* x, float* y, int64_t C, int64_t T, const ecog_sos_plan& p, const SosCoef& coef_in,
                    float* tmp, cudaStream_t st);

// defined in sosfilt_pair.cu
int run_sos_warm_pair(const float* x, float* y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                      const ecog_sos_plan& p, const SosCoef& coef_in, float* tmp, int64_t ldt, double* padbuf,
                      cudaStream_t st);

}  // namespace ecog
