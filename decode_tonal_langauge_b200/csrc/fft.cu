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
// transform; the digit reversal is applied for free when the tile rows are loaded
// (each row is an independent 64-byte segment).  Root-of-unity tables are built by the host
// in float64 and rounded once (decode_tonal_langauge_b200/fftplan.py).
// The repack kernel turns Z into rfft bins, applies resample's truncation / Nyquist rule and
// folds the half spectrum for the inverse real transform, which runs as a conjugated forward
// four-step FFT of N' = num/2 points.  Only the 2/r of the forward spectrum that survives
// the down-sampling is written by pass B.
#include "common.cuh"

namespace ecog {

constexpr int kFftThreads = 256;
constexpr int kW = 8;             // tile width (columns): 64-byte row segments
constexpr int kWp = kW + 1;       // padded pitch (float2)
constexpr int kBigShift = 12;     // two-level twiddle split: e = hi * 4096 + lo

struct AxisDev {
    int n, nstage;
    int radix[16], lprev[16], twstride[16];
    unsigned magic[16];           // ceil(2^32 / lprev): floor(b / lprev) == __umulhi(b, magic) for b < 2^16
};

struct PassParams {
    const float2* in;
    float2* out;
    long long in_ch_stride, out_ch_stride;   // float2 elements per channel
    int n, m;                                // matrix view: n rows (FFT length) x m columns
    const int* perm;
    const float2* tw;
    const float2* tw_hi;
    const float2* tw_lo;
    int conj_in, twiddle, transposed, conj_out;
    float scale;
    int keep_lo, keep_hi;                    // non-transposed store keeps rows q <= keep_lo or q >= keep_hi
    AxisDev ax;
};

__device__ __forceinline__ float2 cmulf(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
__device__ __forceinline__ float2 caddf(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csubf(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 mulmi(float2 a) { return make_float2(a.y, -a.x); }   // * (-i)

__global__ void __launch_bounds__(kFftThreads)
fft_tile_kernel(const PassParams P) {
    extern __shared__ __align__(16) float2 tile[];      // [n][kWp]
    const int tid = threadIdx.x;
    const int n = P.n, m = P.m;
    const int c0 = blockIdx.x * kW;
    const float2* in = P.in + (long long)blockIdx.y * P.in_ch_stride;
    float2* out = P.out + (long long)blockIdx.y * P.out_ch_stride;
    const int cw = m - c0 < kW ? m - c0 : kW;           // valid columns in this tile

    // ---- load: row i (cw contiguous complex) -> tile[perm[i]]
    for (int idx = tid; idx < n * kW; idx += kFftThreads) {
        const int i = idx >> 3, c = idx & (kW - 1);
        float2 v = make_float2(0.f, 0.f);
        if (c < cw) {
            v = in[(long long)i * m + c0 + c];
            if (P.conj_in) v.y = -v.y;
        }
        tile[__ldg(&P.perm[i]) * kWp + c] = v;
    }
    __syncthreads();

    // ---- in-place DIT stages
    for (int s = 0; s < P.ax.nstage; ++s) {
        const int r = P.ax.radix[s], lp = P.ax.lprev[s], tws = P.ax.twstride[s];
        const unsigned magic = P.ax.magic[s];
        const int nbf = n / r;                           // butterflies per column
        for (int w = tid; w < nbf * kW; w += kFftThreads) {
            const int c = w & (kW - 1);
            const int b = w >> 3;
            const int g = lp == 1 ? b : (int)__umulhi((unsigned)b, magic);
            const int j = b - g * lp;
            float2* e0 = tile + (g * lp * r + j) * kWp + c;
            const int step = lp * kWp;
            if (r == 4) {
                float2 a0 = e0[0], a1 = e0[step], a2 = e0[2 * step], a3 = e0[3 * step];
                if (j) {
                    a1 = cmulf(a1, __ldg(&P.tw[j * tws]));
                    a2 = cmulf(a2, __ldg(&P.tw[2 * j * tws]));
                    a3 = cmulf(a3, __ldg(&P.tw[3 * j * tws]));
                }
                float2 t0 = caddf(a0, a2), t1 = csubf(a0, a2), t2 = caddf(a1, a3), t3 = mulmi(csubf(a1, a3));
                e0[0] = caddf(t0, t2); e0[step] = caddf(t1, t3);
                e0[2 * step] = csubf(t0, t2); e0[3 * step] = csubf(t1, t3);
            } else if (r == 2) {
                float2 a0 = e0[0], a1 = e0[step];
                if (j) a1 = cmulf(a1, __ldg(&P.tw[j * tws]));
                e0[0] = caddf(a0, a1); e0[step] = csubf(a0, a1);
            } else if (r == 3) {
                float2 a0 = e0[0], a1 = e0[step], a2 = e0[2 * step];
                if (j) {
                    a1 = cmulf(a1, __ldg(&P.tw[j * tws]));
                    a2 = cmulf(a2, __ldg(&P.tw[2 * j * tws]));
                }
                const float S3 = 0.86602540378443865f;
                float2 t1 = caddf(a1, a2);
                float2 t2 = make_float2(a0.x - 0.5f * t1.x, a0.y - 0.5f * t1.y);
                float2 d = csubf(a1, a2);
                float2 t3 = mulmi(make_float2(S3 * d.x, S3 * d.y));
                e0[0] = caddf(a0, t1); e0[step] = caddf(t2, t3); e0[2 * step] = csubf(t2, t3);
            } else {   // r == 5
                float2 a0 = e0[0], a1 = e0[step], a2 = e0[2 * step], a3 = e0[3 * step], a4 = e0[4 * step];
                if (j) {
                    a1 = cmulf(a1, __ldg(&P.tw[j * tws]));
                    a2 = cmulf(a2, __ldg(&P.tw[2 * j * tws]));
                    a3 = cmulf(a3, __ldg(&P.tw[3 * j * tws]));
                    a4 = cmulf(a4, __ldg(&P.tw[4 * j * tws]));
                }
                const float C1 = 0.30901699437494742f, C2 = -0.80901699437494742f;
                const float S1 = 0.95105651629515357f, S2 = 0.58778525229247313f;
                float2 t1 = caddf(a1, a4), t2 = caddf(a2, a3), t3 = csubf(a1, a4), t4 = csubf(a2, a3);
                float2 m1 = make_float2(a0.x + C1 * t1.x + C2 * t2.x, a0.y + C1 * t1.y + C2 * t2.y);
                float2 m2 = make_float2(a0.x + C2 * t1.x + C1 * t2.x, a0.y + C2 * t1.y + C1 * t2.y);
                float2 n1 = mulmi(make_float2(S1 * t3.x + S2 * t4.x, S1 * t3.y + S2 * t4.y));   // -i n1
                float2 n2 = mulmi(make_float2(S2 * t3.x - S1 * t4.x, S2 * t3.y - S1 * t4.y));   // -i n2
                e0[0] = make_float2(a0.x + t1.x + t2.x, a0.y + t1.y + t2.y);
                e0[step] = caddf(m1, n1); e0[4 * step] = csubf(m1, n1);
                e0[2 * step] = caddf(m2, n2); e0[3 * step] = csubf(m2, n2);
            }
        }
        __syncthreads();
    }

    // ---- store
    if (P.transposed) {
        // out[(c0 + c) * n + q]: contiguous along q; pitch 9 keeps the shared reads conflict free
        for (int c = 0; c < cw; ++c) {
            const long long col = c0 + c;
            for (int q = tid; q < n; q += kFftThreads) {
                float2 v = tile[q * kWp + c];
                if (P.twiddle) {
                    const long long e = (long long)q * col;              // < N
                    const float2 wh = __ldg(&P.tw_hi[e >> kBigShift]);
                    const float2 wl = __ldg(&P.tw_lo[e & ((1 << kBigShift) - 1)]);
                    v = cmulf(v, cmulf(wh, wl));
                }
                out[col * n + q] = v;
            }
        }
    } else {
        for (int idx = tid; idx < n * kW; idx += kFftThreads) {
            const int q = idx >> 3, c = idx & (kW - 1);
            if (c < cw && (q <= P.keep_lo || q >= P.keep_hi)) {
                float2 v = tile[q * kWp + c];
                v.x *= P.scale; v.y *= P.conj_out ? -P.scale : P.scale;
                out[(long long)q * m + c0 + c] = v;
            }
        }
    }
}

// rfft untangle + resample bin rules + fold for the half-length inverse (one thread per k < N')
__global__ void __launch_bounds__(256)
resample_repack_kernel(const float2* __restrict__ Z, float2* __restrict__ G, long long z_stride,
                       long long g_stride, int N, int Nh, long long T, long long num,
                       const float2* __restrict__ twT, const float2* __restrict__ twNum) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= Nh) return;
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
        x.x *= f; x.y *= f;
        if (kk == 0 || kk == Nh) x.y = 0.f;                              // irfft ignores these imaginary parts
        return x;
    };
    const float2 yk = Y(k), yr = Y(Nh - k);
    const float2 ym = make_float2(yr.x, -yr.y);
    const float2 E = make_float2(0.5f * (yk.x + ym.x), 0.5f * (yk.y + ym.y));
    const float2 O = cmulf(make_float2(0.5f * (yk.x - ym.x), 0.5f * (yk.y - ym.y)), __ldg(&twNum[k]));
    // G = E + i O
    G[(long long)blockIdx.y * g_stride + k] = make_float2(E.x - O.y, E.y + O.x);
}

static int fill_axis(const ecog_fft_axis& a, AxisDev& d) {
    if (a.n < 1 || a.nstage < 0 || a.nstage > 16) return fail(ECOG_E_VALUE, "fft axis: bad plan n=%d nstage=%d", a.n, a.nstage);
    d.n = a.n; d.nstage = a.nstage;
    long long L = 1;
    for (int s = 0; s < a.nstage; ++s) {
        const int r = a.radix[s];
        if (r != 2 && r != 3 && r != 4 && r != 5) return fail(ECOG_E_VALUE, "fft axis: radix %d not supported", r);
        d.radix[s] = r; d.lprev[s] = (int)L;
        d.magic[s] = (unsigned)((0x100000000ull + (unsigned long long)L - 1) / (unsigned long long)L);
        L *= r;
        d.twstride[s] = (int)(a.n / L);
    }
    if (L != a.n) return fail(ECOG_E_VALUE, "fft axis: radices multiply to %lld, not n=%d", L, a.n);
    if ((long long)a.n * kWp * (long long)sizeof(float2) > 220 * 1024)
        return fail(ECOG_E_UNSUPPORTED, "fft axis: n=%d does not fit shared memory", a.n);
    if (a.n / 2 >= 65536) return fail(ECOG_E_UNSUPPORTED, "fft axis: n=%d too long", a.n);
    return ECOG_OK;
}

static int launch_pass(PassParams& P, int64_t C, cudaStream_t st, const char* what) {
    const size_t smem = (size_t)P.n * kWp * sizeof(float2);
    ECOG_CUDA(cudaFuncSetAttribute(fft_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((unsigned)ceil_div(P.m, kW), (unsigned)C);
    fft_tile_kernel<<<grid, kFftThreads, smem, st>>>(P);
    return check_launch(what);
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
    P.n = plan->fa.n; P.m = plan->fb.n; P.perm = tb->perm_fa; P.tw = (const float2*)tb->tw_fa;
    P.tw_hi = (const float2*)tb->tw_big_f_hi; P.tw_lo = (const float2*)tb->tw_big_f_lo;
    P.twiddle = plan->fb.n > 1; P.transposed = 1; P.scale = 1.f; P.keep_lo = 1 << 30; P.keep_hi = 0;
    ECOG_TRY(fill_axis(plan->fa, P.ax));
    ECOG_TRY(launch_pass(P, C, st, "fft_fwd_a"));
    if (plan->fb.n > 1) {
        // ---- forward pass B: (fb.n x fa.n), natural-order store of the needed rows only
        memset(&P, 0, sizeof(P));
        P.in = bufA; P.in_ch_stride = N; P.out = bufZ; P.out_ch_stride = N;
        P.n = plan->fb.n; P.m = plan->fa.n; P.perm = tb->perm_fb; P.tw = (const float2*)tb->tw_fb;
        P.scale = 1.f;
        // bins k = p * fa.n + q needed: k <= kneed or k >= N - kneed
        P.keep_lo = (int)(kneed / plan->fa.n);
        P.keep_hi = (int)((N - kneed) / plan->fa.n);
        ECOG_TRY(fill_axis(plan->fb, P.ax));
        ECOG_TRY(launch_pass(P, C, st, "fft_fwd_b"));
    }
    // ---- repack
    {
        dim3 grid((unsigned)ceil_div(Nh, 256), (unsigned)C);
        resample_repack_kernel<<<grid, 256, 0, st>>>(bufZ, bufG, N, Nh, (int)N, (int)Nh, T, num,
                                                     (const float2*)tb->tw_T, (const float2*)tb->tw_num);
        ECOG_TRY(check_launch("resample_repack"));
    }
    // ---- inverse pass A (conjugated input)
    const bool two = plan->ib.n > 1;
    memset(&P, 0, sizeof(P));
    P.in = bufG; P.in_ch_stride = Nh;
    P.out = two ? bufI : reinterpret_cast<float2*>(d_y); P.out_ch_stride = two ? Nh : ldy / 2;
    P.n = plan->ia.n; P.m = plan->ib.n; P.perm = tb->perm_ia; P.tw = (const float2*)tb->tw_ia;
    P.tw_hi = (const float2*)tb->tw_big_i_hi; P.tw_lo = (const float2*)tb->tw_big_i_lo;
    P.conj_in = 1; P.twiddle = two; P.transposed = 1; P.scale = 1.f; P.keep_lo = 1 << 30; P.keep_hi = 0;
    ECOG_TRY(fill_axis(plan->ia, P.ax));
    if (!two) {
        // single pass: the transposed store has m == 1, so it IS natural order; finish here
        P.transposed = 0; P.conj_out = 1; P.scale = (float)(1.0 / (double)Nh);
        return launch_pass(P, C, st, "fft_inv_a");
    }
    ECOG_TRY(launch_pass(P, C, st, "fft_inv_a"));
    memset(&P, 0, sizeof(P));
    P.in = bufI; P.in_ch_stride = Nh; P.out = reinterpret_cast<float2*>(d_y); P.out_ch_stride = ldy / 2;
    P.n = plan->ib.n; P.m = plan->ia.n; P.perm = tb->perm_ib; P.tw = (const float2*)tb->tw_ib;
    P.conj_out = 1; P.scale = (float)(1.0 / (double)Nh); P.keep_lo = 1 << 30; P.keep_hi = 0;
    ECOG_TRY(fill_axis(plan->ib, P.ax));
    return launch_pass(P, C, st, "fft_inv_b");
}
