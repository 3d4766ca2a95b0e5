// K5 whole-row FFT resample (scipy.signal.resample semantics for real input).
//
// A real row of T samples is read as N = T/2 complex points.  The length-N transform is a
// four-step FFT N = n_a * n_b done in two passes of the same tile kernel:
//   pass A : view the row as an (n_a x n_b) matrix, FFT along axis 0 (stride n_b) for a
//            tile of W = 8 adjacent columns held in shared memory, multiply by the
//            four-step twiddle W_N^{q c}, store TRANSPOSED (contiguous along q);
//   pass B : view the intermediate as an (n_b x n_a) matrix, FFT along axis 0 again,
//            store in place layout -> natural order X[q + n_a p].
// The in-shared-memory FFT is an in-place mixed-radix (2,3,4,5) decimation-in-time
// transform; two stages are fused per shared-memory sweep (radix-16/20/25/... "super
// butterflies" held in registers); the digit reversal is applied for free when the tile
// rows are loaded (each row is an independent 32/64-byte segment).  Root-of-unity tables are built by the host
// in float64 and rounded once (decode_tonal_langauge_b200/fftplan.py).
// The repack kernel turns Z into rfft bins, applies resample's truncation / Nyquist rule and
// folds the half spectrum for the inverse real transform, which runs as a conjugated forward
// four-step FFT of N' = num/2 points.  Only the 2/r of the forward spectrum that survives
// the down-sampling is written by pass B.
#include "common.cuh"

namespace ecog {

constexpr int kFftThreads = 384;   // two CTAs per SM; tile + root table in shared memory
constexpr int kBigShift = 12;     // two-level twiddle split: e = hi * 4096 + lo

// One shared-memory sweep = one or two fused DIT stages done in registers.
struct AxisDev {
    int n, npass;
    int ra[16], rb[16];           // radices of the fused stage pair (rb == 1: single stage)
    int lprev[16];                // sub-transform length before the pass
    int tws_a[16], tws_b[16];     // table strides n / L_t, n / L_{t+1}
    unsigned magic[16];           // ceil(2^32 / lprev)
    unsigned magic_nsb[16];       // ceil(2^32 / (n / (ra * rb))): super-butterflies per column
    int nrest;                    // radix stages after the first sweep ...
    int rest[16];                 // ... and their radices (digit order of the super-butterfly index g)
};

struct PassParams {
    const float2* in;
    float2* out;
    long long in_ch_stride, out_ch_stride;   // float2 elements per channel
    int n, m;                                // matrix view: n rows (FFT length) x m columns
    const float2* tw;
    const float2* tw_hi;
    const float2* tw_lo;
    const double2* tw_q;                     // W_N^q, q < n, float64 (four-step twiddle base)
    int conj_in, twiddle, transposed, conj_out;
    int tw_shared;                           // root table staged in shared memory (it fits beside the tile)
    float scale;
    int keep_lo, keep_hi;                    // non-transposed store keeps rows q <= keep_lo or q >= keep_hi
    AxisDev ax;
};

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
#ifndef ECOG_SCALAR_F32
// complex add / sub as ONE packed instruction on the (re, im) register pair (sm_100 add.f32x2 / sub.f32x2)
__device__ __forceinline__ float2 caddf(float2 a, float2 b) {
    float2 r;
    asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 csubf(float2 a, float2 b) {
    float2 r;
    asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; sub.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
#else
__device__ __forceinline__ float2 caddf(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csubf(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
#endif
__device__ __forceinline__ float2 mulmi(float2 a) { return make_float2(a.y, -a.x); }   // * (-i)

// forward R-point DFT on a[0], a[S], a[2S], ... (register array, compile-time indices)
template <int R, int S>
__device__ __forceinline__ void bfly(float2* a) {
    if (R == 2) {
        float2 x0 = a[0], x1 = a[S];
        a[0] = caddf(x0, x1); a[S] = csubf(x0, x1);
    } else if (R == 3) {
        const float S3 = 0.86602540378443865f;
        float2 x0 = a[0], x1 = a[S], x2 = a[2 * S];
        float2 t1 = caddf(x1, x2);
        float2 t2 = make_float2(fmaf(-0.5f, t1.x, x0.x), fmaf(-0.5f, t1.y, x0.y));
        float2 d = csubf(x1, x2);
        float2 t3 = make_float2(S3 * d.y, -S3 * d.x);                       // -i * S3 * d
        a[0] = caddf(x0, t1); a[S] = caddf(t2, t3); a[2 * S] = csubf(t2, t3);
    } else if (R == 4) {
        float2 x0 = a[0], x1 = a[S], x2 = a[2 * S], x3 = a[3 * S];
        float2 t0 = caddf(x0, x2), t1 = csubf(x0, x2), t2 = caddf(x1, x3), t3 = mulmi(csubf(x1, x3));
        a[0] = caddf(t0, t2); a[S] = caddf(t1, t3); a[2 * S] = csubf(t0, t2); a[3 * S] = csubf(t1, t3);
    } else if (R == 5) {
        const float C1 = 0.30901699437494742f, C2 = -0.80901699437494742f;
        const float S1 = 0.95105651629515357f, S2 = 0.58778525229247313f;
        float2 x0 = a[0], x1 = a[S], x2 = a[2 * S], x3 = a[3 * S], x4 = a[4 * S];
        float2 t1 = caddf(x1, x4), t2 = caddf(x2, x3), t3 = csubf(x1, x4), t4 = csubf(x2, x3);
        float2 m1 = make_float2(fmaf(C2, t2.x, fmaf(C1, t1.x, x0.x)), fmaf(C2, t2.y, fmaf(C1, t1.y, x0.y)));
        float2 m2 = make_float2(fmaf(C1, t2.x, fmaf(C2, t1.x, x0.x)), fmaf(C1, t2.y, fmaf(C2, t1.y, x0.y)));
        float2 u1 = make_float2(fmaf(S2, t4.x, S1 * t3.x), fmaf(S2, t4.y, S1 * t3.y));
        float2 u2 = make_float2(fmaf(-S1, t4.x, S2 * t3.x), fmaf(-S1, t4.y, S2 * t3.y));
        float2 n1 = mulmi(u1), n2 = mulmi(u2);
        a[0] = make_float2(x0.x + t1.x + t2.x, x0.y + t1.y + t2.y);
        a[S] = caddf(m1, n1); a[4 * S] = csubf(m1, n1);
        a[2 * S] = caddf(m2, n2); a[3 * S] = csubf(m2, n2);
    }
}

// Two fused DIT stages (radix RA then radix RB) of one column, all in registers.
//   elements  base + qa*lp + qb*Lt   (Lt = lp*RA),  base = g*Lt*RB + j,  j < lp
template <int RA, int RB, int WP>
__device__ __forceinline__ void super_bfly(float2* col, int g, int j, int lp, const float2* __restrict__ tw,
                                           int tws_a, int tws_b) {
    const int Lt = lp * RA;
    float2 a[RA * RB];                       // a[qa * RB + qb]
    float2* e0 = col + (g * Lt * RB + j) * WP;
#pragma unroll
    for (int qb = 0; qb < RB; ++qb)
#pragma unroll
        for (int qa = 0; qa < RA; ++qa) a[qa * RB + qb] = e0[(qa * lp + qb * Lt) * WP];
    if (j) {
#pragma unroll
        for (int qa = 1; qa < RA; ++qa) {
            const float2 w = tw[j * qa * tws_a];
#pragma unroll
            for (int qb = 0; qb < RB; ++qb) a[qa * RB + qb] = cmulf(a[qa * RB + qb], w);
        }
    }
#pragma unroll
    for (int qb = 0; qb < RB; ++qb) bfly<RA, RB>(a + qb);              // over qa, stride RB
    if (RB > 1) {
#pragma unroll
        for (int pa = 0; pa < RA; ++pa) {
            const int jp = j + pa * lp;
            if (jp) {
#pragma unroll
                for (int qb = 1; qb < RB; ++qb)
                    a[pa * RB + qb] = cmulf(a[pa * RB + qb], tw[jp * qb * tws_b]);
            }
            bfly<RB, 1>(a + pa * RB);                                  // over qb, stride 1
        }
    }
#pragma unroll
    for (int pb = 0; pb < RB; ++pb)
#pragma unroll
        for (int pa = 0; pa < RA; ++pa) e0[(pa * lp + pb * Lt) * WP] = a[pa * RB + pb];
}

// First sweep (lp == 1) of one column fed straight from global memory: super-butterfly g owns the
// tile rows g R + qa + qb RA (R = RA RB), i.e. the input rows i0 + qa n/RA + qb n/R, where i0 is g
// with its mixed-radix digits reversed (the digit-reversal permutation of the DIT transform, built
// here instead of being read from a table).  Only the inner twiddles W_{RA RB}^{pa qb} appear (j == 0);
// they come from global memory, the shared root table is not published yet.
template <int RA, int RB, int WP>
__device__ __forceinline__ void super_bfly_first(float2* col, const float2* __restrict__ src, bool ok, bool conj,
                                                 int g, long long i0m, long long stride_a, long long stride_b,
                                                 const float2* __restrict__ twg, int tws_b) {
    float2 a[RA * RB];                       // a[qa * RB + qb]
#pragma unroll
    for (int qb = 0; qb < RB; ++qb)
#pragma unroll
        for (int qa = 0; qa < RA; ++qa) {
            float2 v = make_float2(0.f, 0.f);
            if (ok) v = __ldg(src + i0m + qa * stride_a + qb * stride_b);
            a[qa * RB + qb] = v;
        }
    if (conj) {
#pragma unroll
        for (int e = 0; e < RA * RB; ++e) a[e].y = -a[e].y;
    }
#pragma unroll
    for (int qb = 0; qb < RB; ++qb) bfly<RA, RB>(a + qb);              // over qa, stride RB
    if (RB > 1) {
#pragma unroll
        for (int pa = 0; pa < RA; ++pa) {
            if (pa) {
#pragma unroll
                for (int qb = 1; qb < RB; ++qb) a[pa * RB + qb] = cmulf(a[pa * RB + qb], __ldg(&twg[pa * qb * tws_b]));   // W_{RA RB}^{pa qb}
            }
            bfly<RB, 1>(a + pa * RB);                                  // over qb, stride 1
        }
    }
    float2* e0 = col + (g * RA * RB) * WP;
#pragma unroll
    for (int pb = 0; pb < RB; ++pb)
#pragma unroll
        for (int pa = 0; pa < RA; ++pa) e0[(pa + pb * RA) * WP] = a[pa * RB + pb];
}

template <int W>
__global__ void __launch_bounds__(kFftThreads, 2)
fft_tile_kernel(const PassParams P) {
    constexpr int WP = W + 1;                           // padded pitch (float2)
    constexpr int LOGW = W == 8 ? 3 : 2;
    extern __shared__ __align__(16) float2 tile[];      // [n][WP] | roots of unity [n]
    const int tid = threadIdx.x;
    const int n = P.n, m = P.m;
    // W_n^k: the sweeps read their twiddles from shared memory (the L1 left beside two resident
    // tiles is too small to keep the table), or from global memory for axes too long for that
    const float2* tws = P.tw_shared ? tile + (size_t)n * WP : P.tw;
    float2* tws_w = tile + (size_t)n * WP;
    const int c0 = blockIdx.x * W;
    const float2* in = P.in + (long long)blockIdx.y * P.in_ch_stride;
    float2* out = P.out + (long long)blockIdx.y * P.out_ch_stride;
    const int cw = m - c0 < W ? m - c0 : W;             // valid columns in this tile

    // ---- first sweep, fed from global memory (no separate tile fill): work item -> (column, butterfly g),
    // columns fastest so that the W columns of an input row are one contiguous segment
    if (P.tw_shared) for (int k = tid; k < n; k += kFftThreads) tws_w[k] = __ldg(&P.tw[k]);
    if (P.ax.npass == 0) {                               // n == 1: the transform is the identity
        if (tid < W) {
            float2 v = make_float2(0.f, 0.f);
            if (tid < cw) v = in[c0 + tid];
            if (P.conj_in) v.y = -v.y;
            tile[tid] = v;
        }
    } else {
        const int ra = P.ax.ra[0], rb = P.ax.rb[0];
        const int R0 = ra * rb;
        const int nsb = n / R0;
        const long long stride_a = (long long)(n / ra) * m, stride_b = (long long)(n / R0) * m;
        const int tb = P.ax.tws_b[0];
        const int code = ra * 8 + rb;
        for (int w = tid; w < nsb * W; w += kFftThreads) {
            const int c = w & (W - 1);
            const int g = w >> LOGW;
            int i0 = 0, gg = g;                          // digits of g reversed (Horner over the later radices)
            for (int k = 0; k < P.ax.nrest; ++k) {
                const int r = P.ax.rest[k];
                const int q = gg / r;
                i0 = i0 * r + (gg - q * r);
                gg = q;
            }
            const long long i0m = (long long)i0 * m + c0 + c;
            const bool ok = c < cw;
            float2* col = tile + c;
            const bool cj = P.conj_in != 0;
            switch (code) {
                case 4 * 8 + 4: super_bfly_first<4, 4, WP>(col, in, ok, cj, g, i0m, stride_a, stride_b, P.tw, tb); break;
                case 4 * 8 + 2: super_bfly_first<4, 2, WP>(col, in, ok, cj, g, i0m, stride_a, stride_b, P.tw, tb); break;
                case 4 * 8 + 3: super_bfly_first<4, 3, WP>(col, in, ok, cj, g, i0m, stride_a, stride_b, P.tw, tb); break;
                case 4 * 8 + 5: super_bfly_first<4, 5, WP>(col, in, ok, cj, g, i0m, stride_a, stride_b, P.tw, tb); break;
                case 2 * 8 + 3: super_bfly_first<2, 3, WP>(col, in, ok, cj, g, i0m, stride_a, stride_b, P.tw, tb); break;
                case 2 * 8 + 5: super_bfly_first<2, 5, WP>(col, in, ok, cj, g, i0m, stride_a, stride_b, P.tw, tb); break;
                case 3 * 8 + 3: super_bfly_first<3, 3, WP>(col, in, ok, cj, g, i0m, stride_a, stride_b, P.tw, tb); break;
                case 3 * 8 + 5: super_bfly_first<3, 5, WP>(col, in, ok, cj, g, i0m, stride_a, stride_b, P.tw, tb); break;
                case 5 * 8 + 5: super_bfly_first<5, 5, WP>(col, in, ok, cj, g, i0m, stride_a, stride_b, P.tw, tb); break;
                case 2 * 8 + 1: super_bfly_first<2, 1, WP>(col, in, ok, cj, g, i0m, stride_a, stride_b, P.tw, tb); break;
                case 3 * 8 + 1: super_bfly_first<3, 1, WP>(col, in, ok, cj, g, i0m, stride_a, stride_b, P.tw, tb); break;
                case 4 * 8 + 1: super_bfly_first<4, 1, WP>(col, in, ok, cj, g, i0m, stride_a, stride_b, P.tw, tb); break;
                default:        super_bfly_first<5, 1, WP>(col, in, ok, cj, g, i0m, stride_a, stride_b, P.tw, tb); break;
            }
        }
    }
    __syncthreads();

    // ---- in-place DIT: one shared-memory sweep per fused stage pair
    for (int s = 1; s < P.ax.npass; ++s) {
        const int ra = P.ax.ra[s], rb = P.ax.rb[s], lp = P.ax.lprev[s];
        const int ta = P.ax.tws_a[s], tb = P.ax.tws_b[s];
        const unsigned magic = P.ax.magic[s];
        const int nsb = n / (ra * rb);                   // super-butterflies per column
        const int code = ra * 8 + rb;
        // Work item -> (column, super-butterfly).  First sweep (lp == 1): columns fastest, the W
        // columns of a tile row are one contiguous segment.  Later sweeps: butterflies fastest --
        // consecutive lanes then touch consecutive tile rows of ONE column, and with the odd pitch
        // WP a half-warp's 64-bit accesses fall in 16 different bank pairs (columns-fastest puts two
        // 64-byte rows per half-warp, which overlap in the banks: two-way conflicts on every access).
        const bool col_fast = lp == 1 || nsb == 1;
        const unsigned mnsb = P.ax.magic_nsb[s];
        for (int w = tid; w < nsb * W; w += kFftThreads) {
            int c, b;
            if (col_fast) { c = w & (W - 1); b = w >> LOGW; }
            else { c = (int)__umulhi((unsigned)w, mnsb); b = w - c * nsb; }
            const int g = lp == 1 ? b : (int)__umulhi((unsigned)b, magic);
            const int j = b - g * lp;
            float2* col = tile + c;
            switch (code) {
                case 4 * 8 + 4: super_bfly<4, 4, WP>(col, g, j, lp, tws, ta, tb); break;
                case 4 * 8 + 2: super_bfly<4, 2, WP>(col, g, j, lp, tws, ta, tb); break;
                case 4 * 8 + 3: super_bfly<4, 3, WP>(col, g, j, lp, tws, ta, tb); break;
                case 4 * 8 + 5: super_bfly<4, 5, WP>(col, g, j, lp, tws, ta, tb); break;
                case 2 * 8 + 3: super_bfly<2, 3, WP>(col, g, j, lp, tws, ta, tb); break;
                case 2 * 8 + 5: super_bfly<2, 5, WP>(col, g, j, lp, tws, ta, tb); break;
                case 3 * 8 + 3: super_bfly<3, 3, WP>(col, g, j, lp, tws, ta, tb); break;
                case 3 * 8 + 5: super_bfly<3, 5, WP>(col, g, j, lp, tws, ta, tb); break;
                case 5 * 8 + 5: super_bfly<5, 5, WP>(col, g, j, lp, tws, ta, tb); break;
                case 2 * 8 + 1: super_bfly<2, 1, WP>(col, g, j, lp, tws, ta, tb); break;
                case 3 * 8 + 1: super_bfly<3, 1, WP>(col, g, j, lp, tws, ta, tb); break;
                case 4 * 8 + 1: super_bfly<4, 1, WP>(col, g, j, lp, tws, ta, tb); break;
                default:        super_bfly<5, 1, WP>(col, g, j, lp, tws, ta, tb); break;
            }
        }
        __syncthreads();
    }

    // ---- store
    if (P.transposed) {
        // out[(c0 + c) * n + q]: one thread per q walks the W columns of its tile row (pitch WP: the
        // 64-bit shared reads of consecutive q are conflict free), stores are contiguous along q.
        // Four-step twiddle W_N^{q (c0 + c)} = W_N^{q c0} (W_N^q)^c: ONE two-level table look-up per
        // row and tile, the column powers run in float64 from the float64 root W_N^q.
        for (int q = tid; q < n; q += kFftThreads) {
            float2 v[W];
#pragma unroll
            for (int c = 0; c < W; ++c) v[c] = tile[q * WP + c];
            if (P.twiddle) {
                const long long e = (long long)q * c0;                   // < N
                const float2 wh = __ldg(&P.tw_hi[e >> kBigShift]);
                const float2 wl = __ldg(&P.tw_lo[e & ((1 << kBigShift) - 1)]);
                const double2 w1 = __ldg(&P.tw_q[q]);
                const float2 w0 = cmulf(wh, wl);
                double2 wp = w1;
                v[0] = cmulf(v[0], w0);
#pragma unroll
                for (int c = 1; c < W; ++c) {
                    v[c] = cmulf(v[c], cmulf(w0, make_float2((float)wp.x, (float)wp.y)));
                    wp = make_double2(fma(wp.x, w1.x, -wp.y * w1.y), fma(wp.x, w1.y, wp.y * w1.x));
                }
            }
#pragma unroll
            for (int c = 0; c < W; ++c)
                if (c < cw) out[(long long)(c0 + c) * n + q] = v[c];
        }
    } else {
#pragma unroll 4
        for (int idx = tid; idx < n * W; idx += kFftThreads) {
            const int q = idx >> LOGW, c = idx & (W - 1);
            if (c < cw && (q <= P.keep_lo || q >= P.keep_hi)) {
                float2 v = tile[q * WP + c];
                v.x *= P.scale; v.y *= P.conj_out ? -P.scale : P.scale;
                out[(long long)q * m + c0 + c] = v;
            }
        }
    }
}

// rfft untangle + resample bin rules + fold for the half-length inverse.  One thread per pair
// (k, N' - k), k <= N'/2: both folded bins come from the same four spectrum values,
// G[N' - k] = conj(E) + i conj(O) when G[k] = E + i O (W_num^{-(N'-k)} = -conj(W_num^{-k})).
__global__ void __launch_bounds__(256)
resample_repack_kernel(const float2* __restrict__ Z, float2* __restrict__ G, long long z_stride,
                       long long g_stride, int N, int Nh, long long T, long long num,
                       const float2* __restrict__ twT, const float2* __restrict__ twNum,
                       const float* __restrict__ gain) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > Nh / 2) return;
    const float2* z = Z + (long long)blockIdx.y * z_stride;
    const long long m = num < T ? num : T;
    const int mh = (int)(m / 2);
    const float s = (float)((double)num / (double)T);
    const float nyq = (m % 2 == 0 && num != T) ? (num < T ? 2.0f : 0.5f) : 1.0f;
    auto Y = [&](int kk) -> float2 {
        if (kk > mh) return make_float2(0.f, 0.f);
        const float2 zk = z[kk == N ? 0 : kk];
        const float2 zr = z[kk == 0 ? 0 : N - kk];
        const float2 zm = make_float2(zr.x, -zr.y);                      // conj(Z[N-k])
        const float2 sum = caddf(zk, zm), dif = csubf(zk, zm);
        const float2 wd = cmulf(__ldg(&twT[kk]), dif);
        // X = 0.5 sum - 0.5 i wd
        float2 x = make_float2(0.5f * (sum.x + wd.y), 0.5f * (sum.y - wd.x));
        float f = s * (kk == mh ? nyq : 1.0f);
        if (gain) f *= __ldg(&gain[kk]);                                 // 1 / H1[k] of the FIR pre-decimator
        x.x *= f; x.y *= f;
        if (kk == 0 || kk == Nh) x.y = 0.f;                              // irfft ignores these imaginary parts
        return x;
    };
    const float2 yk = Y(k), yr = Y(Nh - k);
    const float2 ym = make_float2(yr.x, -yr.y);
    const float2 E = make_float2(0.5f * (yk.x + ym.x), 0.5f * (yk.y + ym.y));
    const float2 O = cmulf(make_float2(0.5f * (yk.x - ym.x), 0.5f * (yk.y - ym.y)), __ldg(&twNum[k]));
    // G = E + i O, stored CONJUGATED: the inverse runs as conj(FFT(conj(G))) and the first pass
    // then takes its input as is (asynchronous tile copies)
    float2* g = G + (long long)blockIdx.y * g_stride;
    g[k] = make_float2(E.x - O.y, -(E.y + O.x));
    if (k > 0 && 2 * k != Nh) g[Nh - k] = make_float2(E.x + O.y, E.y - O.x);
}

static bool pair_ok(int ra, int rb) {
    const int code = ra * 8 + rb;
    switch (code) {
        case 4 * 8 + 4: case 4 * 8 + 2: case 4 * 8 + 3: case 4 * 8 + 5: case 2 * 8 + 3:
        case 2 * 8 + 5: case 3 * 8 + 3: case 3 * 8 + 5: case 5 * 8 + 5: return true;
        default: return false;
    }
}

// stages (radix list, DIT order) -> fused passes
static int fill_axis(const ecog_fft_axis& a, AxisDev& d) {
    if (a.n < 1 || a.nstage < 0 || a.nstage > 16) return fail(ECOG_E_VALUE, "fft axis: bad plan n=%d nstage=%d", a.n, a.nstage);
    memset(&d, 0, sizeof(d));
    d.n = a.n;
    long long L = 1;
    int s = 0, np = 0;
    while (s < a.nstage) {
        const int ra = a.radix[s];
        if (ra != 2 && ra != 3 && ra != 4 && ra != 5) return fail(ECOG_E_VALUE, "fft axis: radix %d not supported", ra);
        int rb = 1;
        if (s + 1 < a.nstage && pair_ok(ra, a.radix[s + 1])) rb = a.radix[s + 1];
        d.ra[np] = ra; d.rb[np] = rb; d.lprev[np] = (int)L;
        d.magic[np] = (unsigned)((0x100000000ull + (unsigned long long)L - 1) / (unsigned long long)L);
        {
            const unsigned long long nsb = (unsigned long long)a.n / (unsigned long long)(ra * rb);
            d.magic_nsb[np] = (unsigned)((0x100000000ull + nsb - 1) / nsb);
        }
        d.tws_a[np] = (int)(a.n / (L * ra));
        d.tws_b[np] = (int)(a.n / (L * ra * rb));
        L *= (long long)ra * rb;
        s += rb > 1 ? 2 : 1;
        ++np;
    }
    d.npass = np;
    {
        const int used = d.npass ? (d.rb[0] > 1 ? 2 : 1) : 0;
        d.nrest = a.nstage - used;
        for (int k = 0; k < d.nrest; ++k) d.rest[k] = a.radix[used + k];
    }
    if (L != a.n) return fail(ECOG_E_VALUE, "fft axis: radices multiply to %lld, not n=%d", L, a.n);
    if ((long long)a.n * 5 * (long long)sizeof(float2) > 220 * 1024)
        return fail(ECOG_E_UNSUPPORTED, "fft axis: n=%d does not fit shared memory", a.n);
    if (a.n / 2 >= 65536) return fail(ECOG_E_UNSUPPORTED, "fft axis: n=%d too long", a.n);
    return ECOG_OK;
}

static int launch_pass(PassParams& P, int64_t C, cudaStream_t st, const char* what) {
    // 8-column tiles (64-byte segments) when two CTAs still fit an SM, else 4-column tiles;
    // shared memory = padded tile + the axis' root table (when that still fits)
    const size_t tw_bytes = (size_t)P.n * sizeof(float2);
    const size_t tile8 = (size_t)P.n * 9 * sizeof(float2), tile4 = (size_t)P.n * 5 * sizeof(float2);
    const size_t two = 110 * 1024, one = 220 * 1024;
    const bool use8 = tile8 + tw_bytes <= two || (tile4 + tw_bytes > two && tile8 + tw_bytes <= one);
    const size_t tile_bytes = use8 ? tile8 : tile4;
    P.tw_shared = tile_bytes + tw_bytes <= one ? 1 : 0;
    const size_t smem = tile_bytes + (P.tw_shared ? tw_bytes : 0);
    if (use8) {
        ECOG_TRY((smem_attr<fft_tile_kernel<8>>(smem)));
        dim3 grid((unsigned)ceil_div(P.m, 8), (unsigned)C);
        fft_tile_kernel<8><<<grid, kFftThreads, smem, st>>>(P);
    } else {
        ECOG_TRY((smem_attr<fft_tile_kernel<4>>(smem)));
        dim3 grid((unsigned)ceil_div(P.m, 4), (unsigned)C);
        fft_tile_kernel<4><<<grid, kFftThreads, smem, st>>>(P);
    }
    return check_launch(what);
}

// out[c, i] = (i < in_len ? in[c, i] * table[i] : 0)  for i < out_len; real or complex input,
// complex or real-part output.  The glue of the chirp-z path (zero padding + chirp / spectrum
// multiplications); HBM-bound streaming.
template <bool IN_CPLX, bool OUT_CPLX>
__global__ void __launch_bounds__(256)
cplx_modulate_kernel(const float* __restrict__ in, int64_t in_len, int64_t ld_in,
                     const float2* __restrict__ table, float* __restrict__ out, int64_t out_len, int64_t ld_out) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= out_len) return;
    const int64_t c = blockIdx.y;
    float2 r = make_float2(0.f, 0.f);
    if (i < in_len) {                                   // the table (if any) has in_len entries
        float2 v = make_float2(0.f, 0.f);
        if (IN_CPLX) v = reinterpret_cast<const float2*>(in)[c * ld_in + i];
        else v.x = in[c * ld_in + i];
        const float2 w = table ? __ldg(&table[i]) : make_float2(1.f, 0.f);
        r = IN_CPLX ? cmulf(v, w) : make_float2(v.x * w.x, v.x * w.y);
    }
    if (OUT_CPLX) reinterpret_cast<float2*>(out)[c * ld_out + i] = r;
    else out[c * ld_out + i] = r.x;
}

// acc[c, i] (+)= scale * |z[c, i]|  (or scale * Re z[c, i]): band accumulation of the whole-record
// Gaussian-Hilbert path
template <bool ENV>
__global__ void __launch_bounds__(256)
cplx_absacc_kernel(const float2* __restrict__ z, int64_t ld_z, float* __restrict__ acc, int64_t ld_acc,
                   int64_t len, float scale, int accumulate) {
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= len) return;
    const int64_t c = blockIdx.y;
    const float2 v = z[c * ld_z + i];
    float r = ENV ? sqrtf(fmaf(v.x, v.x, v.y * v.y)) : v.x;
    r *= scale;
    if (accumulate) r += acc[c * ld_acc + i];
    acc[c * ld_acc + i] = r;
}

// generic complex four-step FFT of C rows of N = a.n * b.n points (forward, or inverse as
// conj -> forward -> conj), natural order in and out
static int run_c2c(const float2* in, float2* out, int64_t C, int64_t ld_in, int64_t ld_out,
                   const ecog_fft_axis& fa, const ecog_fft_axis& fb, const ecog_fft_tables* tb,
                   int inverse, float scale, float2* tmp, cudaStream_t st) {
    const int64_t N = (int64_t)fa.n * fb.n;
    PassParams P;
    memset(&P, 0, sizeof(P));
    P.in = in; P.in_ch_stride = ld_in;
    P.n = fa.n; P.m = fb.n; P.tw = (const float2*)tb->tw_a;
    P.tw_hi = (const float2*)tb->tw_big_hi; P.tw_lo = (const float2*)tb->tw_big_lo;
    P.tw_q = (const double2*)tb->tw_q;
    P.conj_in = inverse ? 1 : 0; P.scale = 1.f; P.keep_lo = 1 << 30; P.keep_hi = 0;
    ECOG_TRY(fill_axis(fa, P.ax));
    if (fb.n == 1) {     // one pass: natural-order store, finish here
        P.out = out; P.out_ch_stride = ld_out; P.transposed = 0; P.twiddle = 0;
        P.conj_out = inverse ? 1 : 0; P.scale = scale;
        return launch_pass(P, C, st, "fft_c2c");
    }
    P.out = tmp; P.out_ch_stride = N; P.transposed = 1; P.twiddle = 1;
    ECOG_TRY(launch_pass(P, C, st, "fft_c2c_a"));
    memset(&P, 0, sizeof(P));
    P.in = tmp; P.in_ch_stride = N; P.out = out; P.out_ch_stride = ld_out;
    P.n = fb.n; P.m = fa.n; P.tw = (const float2*)tb->tw_b;
    P.conj_out = inverse ? 1 : 0; P.scale = scale; P.keep_lo = 1 << 30; P.keep_hi = 0;
    ECOG_TRY(fill_axis(fb, P.ax));
    return launch_pass(P, C, st, "fft_c2c_b");
}

static size_t align256(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace ecog

using namespace ecog;

extern "C" size_t ecog_resample_workspace(const ecog_resample_plan* plan, int64_t C) {
    if (!plan) return 0;
    const size_t N = (size_t)plan->T / 2, Nh = (size_t)plan->num / 2;
    // forward intermediate + spectrum Z (N each), folded spectrum G + inverse intermediate (N' each)
    return 2 * align256((size_t)C * N * sizeof(float2)) + 2 * align256((size_t)C * Nh * sizeof(float2));
}

extern "C" int ecog_fft_resample(const float* d_x, float* d_y, int64_t C, int64_t ldx, int64_t ldy,
                                 const ecog_resample_plan* plan, const ecog_resample_tables* tb,
                                 void* d_workspace, size_t workspace_bytes, ecog_stream_t stream) {
    if (!plan || !tb) return fail(ECOG_E_VALUE, "ecog_fft_resample: null plan");
    const int64_t T = plan->T, num = plan->num;
    if (C <= 0 || C > 65535 || T < 2 || num < 2 || ldx < T || ldy < num) return fail(ECOG_E_VALUE, "ecog_fft_resample: bad shape");
    if (T % 2 || num % 2 || ldx % 2 || ldy % 2)
        return fail(ECOG_E_UNSUPPORTED, "ecog_fft_resample: odd lengths / strides are not implemented (T=%lld num=%lld)",
                    (long long)T, (long long)num);
    if ((reinterpret_cast<uintptr_t>(d_x) & 7u) || (reinterpret_cast<uintptr_t>(d_y) & 7u))
        return fail(ECOG_E_VALUE, "ecog_fft_resample: rows must be 8-byte aligned");
    const int64_t N = T / 2, Nh = num / 2;
    if ((int64_t)plan->fa.n * plan->fb.n != N || (int64_t)plan->ia.n * plan->ib.n != Nh)
        return fail(ECOG_E_VALUE, "ecog_fft_resample: plan factors do not match the lengths");
    if (N >= (1ll << 31) / 2) return fail(ECOG_E_UNSUPPORTED, "ecog_fft_resample: row too long");
    if (workspace_bytes < ecog_resample_workspace(plan, C))
        return fail(ECOG_E_WORKSPACE, "ecog_fft_resample: workspace %zu < %zu", workspace_bytes, ecog_resample_workspace(plan, C));
    cudaStream_t st = (cudaStream_t)stream;
    char* ws = (char*)d_workspace;
    float2* bufA = (float2*)ws;                 ws += align256((size_t)C * N * sizeof(float2));
    float2* bufZ = (float2*)ws;                 ws += align256((size_t)C * N * sizeof(float2));
    float2* bufG = (float2*)ws;                 ws += align256((size_t)C * Nh * sizeof(float2));
    float2* bufI = (float2*)ws;

    const int64_t m = num < T ? num : T;
    const int64_t kneed = (m / 2 < Nh ? m / 2 : Nh);       // highest forward bin the repack reads

    PassParams P;
    // ---- forward pass A: (fa.n x fb.n), twiddle, transposed
    memset(&P, 0, sizeof(P));
    P.in = reinterpret_cast<const float2*>(d_x); P.in_ch_stride = ldx / 2;
    P.out = plan->fb.n > 1 ? bufA : bufZ; P.out_ch_stride = N;
    P.n = plan->fa.n; P.m = plan->fb.n; P.tw = (const float2*)tb->tw_fa;
    P.tw_hi = (const float2*)tb->tw_big_f_hi; P.tw_lo = (const float2*)tb->tw_big_f_lo;
    P.tw_q = (const double2*)tb->tw_q_f;
    P.twiddle = plan->fb.n > 1; P.transposed = 1; P.scale = 1.f; P.keep_lo = 1 << 30; P.keep_hi = 0;
    ECOG_TRY(fill_axis(plan->fa, P.ax));
    ECOG_TRY(launch_pass(P, C, st, "fft_fwd_a"));
    if (plan->fb.n > 1) {
        // ---- forward pass B: (fb.n x fa.n), natural-order store of the needed rows only
        memset(&P, 0, sizeof(P));
        P.in = bufA; P.in_ch_stride = N; P.out = bufZ; P.out_ch_stride = N;
        P.n = plan->fb.n; P.m = plan->fa.n; P.tw = (const float2*)tb->tw_fb;
        P.scale = 1.f;
        // bins k = p * fa.n + q needed: k <= kneed or k >= N - kneed
        P.keep_lo = (int)(kneed / plan->fa.n);
        P.keep_hi = (int)((N - kneed) / plan->fa.n);
        ECOG_TRY(fill_axis(plan->fb, P.ax));
        ECOG_TRY(launch_pass(P, C, st, "fft_fwd_b"));
    }
    // ---- repack
    {
        dim3 grid((unsigned)ceil_div(Nh / 2 + 1, 256), (unsigned)C);
        resample_repack_kernel<<<grid, 256, 0, st>>>(bufZ, bufG, N, Nh, (int)N, (int)Nh, T, num,
                                                     (const float2*)tb->tw_T, (const float2*)tb->tw_num,
                                                     tb->bin_gain);
        ECOG_TRY(check_launch("resample_repack"));
    }
    // ---- inverse pass A (conjugated input)
    const bool two = plan->ib.n > 1;
    memset(&P, 0, sizeof(P));
    P.in = bufG; P.in_ch_stride = Nh;
    P.out = two ? bufI : reinterpret_cast<float2*>(d_y); P.out_ch_stride = two ? Nh : ldy / 2;
    P.n = plan->ia.n; P.m = plan->ib.n; P.tw = (const float2*)tb->tw_ia;
    P.tw_hi = (const float2*)tb->tw_big_i_hi; P.tw_lo = (const float2*)tb->tw_big_i_lo;
    P.tw_q = (const double2*)tb->tw_q_i;
    P.conj_in = 0; P.twiddle = two; P.transposed = 1; P.scale = 1.f; P.keep_lo = 1 << 30; P.keep_hi = 0;
    ECOG_TRY(fill_axis(plan->ia, P.ax));
    if (!two) {
        // single pass: the transposed store has m == 1, so it IS natural order; finish here
        P.transposed = 0; P.conj_out = 1; P.scale = (float)(1.0 / (double)Nh);
        return launch_pass(P, C, st, "fft_inv_a");
    }
    ECOG_TRY(launch_pass(P, C, st, "fft_inv_a"));
    memset(&P, 0, sizeof(P));
    P.in = bufI; P.in_ch_stride = Nh; P.out = reinterpret_cast<float2*>(d_y); P.out_ch_stride = ldy / 2;
    P.n = plan->ib.n; P.m = plan->ia.n; P.tw = (const float2*)tb->tw_ib;
    P.conj_out = 1; P.scale = (float)(1.0 / (double)Nh); P.keep_lo = 1 << 30; P.keep_hi = 0;
    ECOG_TRY(fill_axis(plan->ib, P.ax));
    return launch_pass(P, C, st, "fft_inv_b");
}

extern "C" size_t ecog_fft_c2c_workspace(const ecog_fft_axis* fa, const ecog_fft_axis* fb, int64_t C) {
    if (!fa || !fb || fb->n <= 1) return 256;
    return align256((size_t)C * fa->n * fb->n * sizeof(float2));
}

extern "C" int ecog_fft_c2c(const float* d_in, float* d_out, int64_t C, int64_t ld_in, int64_t ld_out,
                            const ecog_fft_axis* fa, const ecog_fft_axis* fb, const ecog_fft_tables* tables,
                            int32_t inverse, float scale, void* d_workspace, size_t workspace_bytes,
                            ecog_stream_t stream) {
    if (!fa || !fb || !tables) return fail(ECOG_E_VALUE, "ecog_fft_c2c: null plan");
    const int64_t N = (int64_t)fa->n * fb->n;
    if (C <= 0 || C > 65535 || N < 1 || ld_in < N || ld_out < N) return fail(ECOG_E_VALUE, "ecog_fft_c2c: bad shape");
    if ((reinterpret_cast<uintptr_t>(d_in) & 7u) || (reinterpret_cast<uintptr_t>(d_out) & 7u))
        return fail(ECOG_E_VALUE, "ecog_fft_c2c: rows must be 8-byte aligned");
    if (workspace_bytes < ecog_fft_c2c_workspace(fa, fb, C))
        return fail(ECOG_E_WORKSPACE, "ecog_fft_c2c: workspace %zu < %zu", workspace_bytes, ecog_fft_c2c_workspace(fa, fb, C));
    return run_c2c(reinterpret_cast<const float2*>(d_in), reinterpret_cast<float2*>(d_out), C, ld_in, ld_out,
                   *fa, *fb, tables, inverse, scale, reinterpret_cast<float2*>(d_workspace), (cudaStream_t)stream);
}

extern "C" int ecog_cplx_modulate(const float* d_in, int32_t in_is_complex, int64_t in_len, int64_t ld_in,
                                  const float* d_table, float* d_out, int32_t out_is_complex, int64_t out_len,
                                  int64_t ld_out, int64_t C, ecog_stream_t stream) {
    if (C <= 0 || C > 65535 || out_len <= 0 || in_len < 0 || ld_out < out_len || ld_in < (in_len < out_len ? in_len : out_len))
        return fail(ECOG_E_VALUE, "ecog_cplx_modulate: bad shape");
    if (!d_in || !d_out) return fail(ECOG_E_VALUE, "ecog_cplx_modulate: null pointer");
    dim3 grid((unsigned)ceil_div(out_len, 256), (unsigned)C);
    cudaStream_t st = (cudaStream_t)stream;
    const float2* tb = reinterpret_cast<const float2*>(d_table);
    if (in_is_complex && out_is_complex) cplx_modulate_kernel<true, true><<<grid, 256, 0, st>>>(d_in, in_len, ld_in, tb, d_out, out_len, ld_out);
    else if (in_is_complex) cplx_modulate_kernel<true, false><<<grid, 256, 0, st>>>(d_in, in_len, ld_in, tb, d_out, out_len, ld_out);
    else if (out_is_complex) cplx_modulate_kernel<false, true><<<grid, 256, 0, st>>>(d_in, in_len, ld_in, tb, d_out, out_len, ld_out);
    else cplx_modulate_kernel<false, false><<<grid, 256, 0, st>>>(d_in, in_len, ld_in, tb, d_out, out_len, ld_out);
    return check_launch("cplx_modulate");
}

extern "C" int ecog_cplx_abs_accumulate(const float* d_z, int64_t ld_z, float* d_acc, int64_t ld_acc, int64_t len,
                                        int64_t C, int32_t envelope, float scale, int32_t accumulate,
                                        ecog_stream_t stream) {
    if (C <= 0 || C > 65535 || len <= 0 || ld_z < len || ld_acc < len || !d_z || !d_acc)
        return fail(ECOG_E_VALUE, "ecog_cplx_abs_accumulate: bad shape");
    dim3 grid((unsigned)ceil_div(len, 256), (unsigned)C);
    cudaStream_t st = (cudaStream_t)stream;
    const float2* z = reinterpret_cast<const float2*>(d_z);
    if (envelope) cplx_absacc_kernel<true><<<grid, 256, 0, st>>>(z, ld_z, d_acc, ld_acc, len, scale, accumulate);
    else cplx_absacc_kernel<false><<<grid, 256, 0, st>>>(z, ld_z, d_acc, ld_acc, len, scale, accumulate);
    return check_launch("cplx_abs_accumulate");
}
