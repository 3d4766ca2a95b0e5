// K3 IIR biquad cascade: zero-phase (filtfilt, odd padding) and causal (sosfilt).
//
// Block-parallel linear recurrence.  A row of T samples is cut into chunks of L
// samples; one thread owns one chunk and runs the float64 DF2T cascade over it.
//   1. tail  : every full chunk is run from a ZERO state over its last `tail` samples
//              (the part of the zero-state response that still reaches the chunk end
//              in float64); the end state g_k is the affine term of the chunk map.
//   2. scan  : per row, S_{k+1} = A^L S_k + g_k carries the true state across chunks
//              (A^L is built by the host); S_0 comes from the filtfilt start-up:
//              zi * ext[0] pushed through the odd-extension pad.
//   3. main  : every chunk is re-run from its true start state and writes float32.
// The backward sweep repeats 1-3 on the forward result, in place, time reversed.
//
// Data movement: 512 chunks per CTA, two CTAs per SM; 16-sample time tiles of all 512
// chunks are staged through shared memory with a double-buffered cp.async pipeline
// (64 B per chunk per stage, coalesced), each thread reads its own tile row with
// conflict-free 128-bit shared loads (row pitch 20 words), overwrites it with the results,
// and the tile is stored with 128-bit coalesced writes.  Arithmetic is FP64-pipe bound
// (5 DFMA per biquad per sample + 2 conversions), see DESIGN.md.
#include "sos_common.cuh"

namespace ecog {

template <int NSEC, bool REV, bool WRITE, bool VEC>
__global__ void __launch_bounds__(kSosThreads, 2)
sos_chunk_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t C, int64_t T,
                 int64_t ldx, int64_t ldy, int L, int tail, int nChunks, int padlen,
                 SosCoef coef, double* __restrict__ state, double* __restrict__ gbuf,
                 double* __restrict__ padbuf) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* tiles = reinterpret_cast<float*>(smem_raw);                         // [kRing][512][kPitch]
    int64_t* gbase = reinterpret_cast<int64_t*>(tiles + (size_t)kRing * kSosThreads * kPitch);  // [512] row
    int* cstart = reinterpret_cast<int*>(gbase + kSosThreads);                 // [512] chunk edge a (fwd) / b (rev)
    int* clen = cstart + kSosThreads;                                          // [512]

    const int tid = threadIdx.x;
    const int64_t q = (int64_t)blockIdx.x * kSosThreads + tid;
    const int64_t items = C * nChunks;
    const bool valid = q < items;
    const int64_t row = valid ? q / nChunks : 0;
    const int k = valid ? (int)(q - row * nChunks) : 0;
    int64_t a, b;
    if (!REV) { a = (int64_t)k * L; b = a + L < T ? a + L : T; }
    else      { b = T - (int64_t)k * L; a = b - L > 0 ? b - L : 0; }
    const bool need = valid && (WRITE || k < nChunks - 1);
    gbase[tid] = row;
    cstart[tid] = (int)(REV ? b : a);
    clen[tid] = need ? (int)(b - a) : 0;
    __syncthreads();

    double c[NSEC][5], s[NSEC][2];
#pragma unroll
    for (int j = 0; j < NSEC; ++j) {
#pragma unroll
        for (int i = 0; i < 5; ++i) c[j][i] = coef.c[j][i];
        s[j][0] = 0.0; s[j][1] = 0.0;
    }
    if (WRITE && valid) {
        const double* sp = state + q * (2 * NSEC);
#pragma unroll
        for (int j = 0; j < NSEC; ++j) { s[j][0] = sp[2 * j]; s[j][1] = sp[2 * j + 1]; }
    }

    const int nStages = L / kSub;
    const int first = WRITE ? 0 : (L - tail) / kSub;
    const float* src = WRITE && REV ? y : x;          // backward main pass runs in place on y
    const int64_t lds = WRITE && REV ? ldy : ldx;

    // cooperative stage loader: VEC -> 4 pieces of 16 B per row, else 16 pieces of 4 B
    auto issue = [&](int stage) {
        if (stage < nStages) {
            float* dst = tiles + (size_t)(stage % kRing) * kSosThreads * kPitch;
            if (VEC) {
                for (int i = tid; i < kSosThreads * 4; i += kSosThreads) {
                    const int r = i >> 2, p = i & 3;
                    const int len = clen[r];
                    int64_t off; bool ok;
                    if (!REV) { int e = stage * kSub + 4 * p; ok = e < len; off = (int64_t)cstart[r] + e; }
                    else { int e = (stage + 1) * kSub - 4 * p; ok = e <= len; off = (int64_t)cstart[r] - e; }
                    const float* g = src + gbase[r] * lds + (ok ? off : 0);
                    cp_async16_zfill(dst + r * kPitch + 4 * p, g, ok);
                }
            } else {
                for (int i = tid; i < kSosThreads * kSub; i += kSosThreads) {
                    const int r = i / kSub, p = i - r * kSub;
                    const int len = clen[r];
                    int64_t off; bool ok;
                    if (!REV) { int e = stage * kSub + p; ok = e < len; off = (int64_t)cstart[r] + e; }
                    else { int e = (stage + 1) * kSub - p; ok = e <= len; off = (int64_t)cstart[r] - e; }
                    const float* g = src + gbase[r] * lds + (ok ? off : 0);
                    cp_async4_zfill(dst + r * kPitch + p, g, ok);
                }
            }
        }
        cp_async_commit();
    };

    issue(first);
    const int mylen = clen[tid];
    for (int st = first; st < nStages; ++st) {
        cp_async_wait<0>();                 // stage st has landed
        __syncthreads();                    // ... for everyone; the other buffer's stores are done
        issue(st + 1);                      // refill the other buffer while this one is processed
        float* mine = tiles + (size_t)(st % kRing) * kSosThreads * kPitch + tid * kPitch;
        float4 xin[kSub / 4];
#pragma unroll
        for (int v = 0; v < kSub / 4; ++v) {
            if (!REV) {
                xin[v] = *reinterpret_cast<const float4*>(mine + 4 * v);
            } else {
                float4 t4 = *reinterpret_cast<const float4*>(mine + (kSub - 4 - 4 * v));
                xin[v] = make_float4(t4.w, t4.z, t4.y, t4.x);
            }
        }
#pragma unroll
        for (int v = 0; v < kSub / 4; ++v) {
            float4 yv;
            const float4 xv = xin[v];
            const int base = st * kSub + 4 * v;        // logical sample index within the chunk
            if (base + 4 <= mylen) {
                yv.x = (float)sos_step<NSEC>((double)xv.x, c, s);
                yv.y = (float)sos_step<NSEC>((double)xv.y, c, s);
                yv.z = (float)sos_step<NSEC>((double)xv.z, c, s);
                yv.w = (float)sos_step<NSEC>((double)xv.w, c, s);
            } else {   // ragged chunk end: samples past it are zero filled; freeze the state there
                yv = make_float4(0.f, 0.f, 0.f, 0.f);
                if (base + 0 < mylen) yv.x = (float)sos_step<NSEC>((double)xv.x, c, s);
                if (base + 1 < mylen) yv.y = (float)sos_step<NSEC>((double)xv.y, c, s);
                if (base + 2 < mylen) yv.z = (float)sos_step<NSEC>((double)xv.z, c, s);
            }
            if (WRITE) {     // results overwrite this thread's own tile row
                if (!REV) *reinterpret_cast<float4*>(mine + 4 * v) = yv;
                else *reinterpret_cast<float4*>(mine + (kSub - 4 - 4 * v)) = make_float4(yv.w, yv.z, yv.y, yv.x);
            }
        }
        if (WRITE) {
            __syncthreads();
            const float* out_tile = tiles + (size_t)(st % kRing) * kSosThreads * kPitch;
            if (VEC) {
                for (int i = tid; i < kSosThreads * 4; i += kSosThreads) {
                    const int r = i >> 2, p = i & 3;
                    const int len = clen[r];
                    int64_t off; bool ok;
                    if (!REV) { int e = st * kSub + 4 * p; ok = e < len; off = (int64_t)cstart[r] + e; }
                    else { int e = (st + 1) * kSub - 4 * p; ok = e <= len; off = (int64_t)cstart[r] - e; }
                    if (ok) {
                        float4 v4 = *reinterpret_cast<const float4*>(out_tile + r * kPitch + 4 * p);
                        *reinterpret_cast<float4*>(y + gbase[r] * ldy + off) = v4;
                    }
                }
            } else {
                for (int i = tid; i < kSosThreads * kSub; i += kSosThreads) {
                    const int r = i / kSub, p = i - r * kSub;
                    const int len = clen[r];
                    int64_t off; bool ok;
                    if (!REV) { int e = st * kSub + p; ok = e < len; off = (int64_t)cstart[r] + e; }
                    else { int e = (st + 1) * kSub - p; ok = e <= len; off = (int64_t)cstart[r] - e; }
                    if (ok) y[gbase[r] * ldy + off] = out_tile[r * kPitch + p];
                }
            }
        }
    }
    cp_async_wait<0>();

    if (!WRITE) {
        if (need) {   // end state of chunk k: the affine term g_k of the chunk map
            double* sp = gbuf + q * (2 * NSEC);
#pragma unroll
            for (int j = 0; j < NSEC; ++j) { sp[2 * j] = s[j][0]; sp[2 * j + 1] = s[j][1]; }
        }
        return;
    }
    // forward main pass: the thread that owns a row's last chunk runs on through the right
    // odd-extension pad ext[T+i] = 2 x[T-1] - x[T-2-i] (float32, like scipy's odd_ext) and keeps
    // the filtered pad in float64 for the backward start-up.
    if (!REV && padlen > 0 && valid && k == nChunks - 1) {
        const float* xr = x + row * ldx;
        const float xe = xr[T - 1];
        double* pb = padbuf + row * padlen;
        for (int i = 0; i < padlen; ++i) {
            const float e = 2.0f * xe - xr[T - 2 - i];
            pb[i] = sos_step<NSEC>((double)e, c, s);
        }
    }
}

// One warp per row: start-up state, then the sequential carry S_{k+1} = M S_k + g_k over chunks.
// The g_k vectors are fetched 32 chunks at a time with one coalesced load per lane (double
// buffered through shared memory), every lane then runs the same short recurrence out of shared
// memory (no global latency on the dependency chain) and the 32 start states are written back
// coalesced.
constexpr int kScanWarps = 2;

template <int NSEC, bool REV>
__global__ void __launch_bounds__(kScanWarps * 32)
sos_scan_kernel(const float* __restrict__ x, int64_t C, int64_t T, int64_t ldx,
                int nChunks, int padlen, int zero_phase, SosCoef coef, SosMatrix M,
                double* __restrict__ state, const double* __restrict__ gbuf,
                const double* __restrict__ padbuf) {
    constexpr int NS = 2 * NSEC;
    __shared__ double gs[kScanWarps][2][32][NS + 1];
    __shared__ double ss[kScanWarps][32][NS + 1];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.x * kScanWarps + wib;
    if (row >= C) return;                                   // whole warp exits together
    double c[NSEC][5], s[NSEC][2];
#pragma unroll
    for (int j = 0; j < NSEC; ++j) {
#pragma unroll
        for (int i = 0; i < 5; ++i) c[j][i] = coef.c[j][i];
        s[j][0] = 0.0; s[j][1] = 0.0;
    }
    if (zero_phase) {      // every lane computes the same start-up (uniform, broadcast loads)
        if (!REV) {
            const float* xr = x + row * ldx;
            const float x0 = xr[0];
            const float e0 = 2.0f * x0 - xr[padlen];
#pragma unroll
            for (int j = 0; j < NSEC; ++j) { s[j][0] = coef.zi[j][0] * (double)e0; s[j][1] = coef.zi[j][1] * (double)e0; }
            for (int i = 0; i < padlen; ++i) {
                const float e = 2.0f * x0 - xr[padlen - i];
                (void)sos_step<NSEC>((double)e, c, s);
            }
        } else {
            const double* pb = padbuf + row * padlen;
            const double y0 = pb[padlen - 1];
#pragma unroll
            for (int j = 0; j < NSEC; ++j) { s[j][0] = coef.zi[j][0] * y0; s[j][1] = coef.zi[j][1] * y0; }
            for (int i = padlen - 1; i >= 0; --i) (void)sos_step<NSEC>(pb[i], c, s);
        }
    }
    // matrix-vector recurrence spread over the warp: P lanes per output component, each lane owns
    // JP columns of M in registers; partial sums meet with xor shuffles, the new state is
    // re-distributed with one shuffle per owned column.
    constexpr int P = NS >= 8 ? (NS > 8 ? 2 : 4) : (NS >= 4 ? 4 : 2);
    constexpr int JP = (NS + P - 1) / P;
    const int oi = lane / P < NS ? lane / P : NS - 1;     // output component of this lane (clamped for idle lanes)
    const int jg = lane % P;
    double mrow[JP], vj[JP];
#pragma unroll
    for (int u = 0; u < JP; ++u) {
        const int j = jg * JP + u;
        mrow[u] = j < NS ? M.m[oi][j] : 0.0;
        const int jj = j < NS ? j : 0;
        vj[u] = (jj & 1) ? s[jj >> 1][1] : s[jj >> 1][0];
    }
    double vi = (oi & 1) ? s[oi >> 1][1] : s[oi >> 1][0];  // current state component oi
    double* sp = state + row * nChunks * NS;
    const double* gp = gbuf + row * nChunks * NS;
    const int nG = nChunks - 1;                             // g_0 .. g_{nChunks-2}
    auto fetch = [&](int batch, int buf) {                  // lane -> g of chunk batch*32 + lane
        const int k = batch * 32 + lane;
        if (k < nG) {
#pragma unroll
            for (int i = 0; i < NS; ++i) gs[wib][buf][lane][i] = gp[(int64_t)k * NS + i];
        }
    };
    fetch(0, 0);
    for (int batch = 0; batch * 32 < nChunks; ++batch) {
        __syncwarp();
        fetch(batch + 1, (batch + 1) & 1);                  // loads in flight during the recurrence
        const int k0 = batch * 32;
        const int cnt = nChunks - k0 < 32 ? nChunks - k0 : 32;
        for (int t = 0; t < cnt; ++t) {
            if (jg == 0 && lane / P < NS) ss[wib][t][oi] = vi;
            if (k0 + t + 1 < nChunks) {
                double acc = jg == 0 ? gs[wib][batch & 1][t][oi] : 0.0;
#pragma unroll
                for (int u = 0; u < JP; ++u) acc = fma(mrow[u], vj[u], acc);
#pragma unroll
                for (int o = P / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                vi = acc;
#pragma unroll
                for (int u = 0; u < JP; ++u) {
                    const int j = jg * JP + u;
                    vj[u] = __shfl_sync(0xffffffffu, acc, (j < NS ? j : 0) * P);
                }
            }
        }
        __syncwarp();
        if (lane < cnt) {                                   // coalesced write-back of 32 start states
#pragma unroll
            for (int i = 0; i < NS; ++i) sp[(int64_t)(k0 + lane) * NS + i] = ss[wib][lane][i];
        }
    }
}

template <int NSEC, bool REV, int NT, int NUM>
static int launch_warm(const float* x, float* y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                       const ecog_sos_plan& p, int nChunks, const SosCoef& coef, double gain, double* padbuf, bool vec,
                       cudaStream_t st) {
    const size_t smem = ((size_t)kWarmRing * NT * kWPitch) * sizeof(float) + (size_t)NT * (2 * sizeof(int64_t) + sizeof(int2));
    const unsigned grid = (unsigned)ceil_div(C * nChunks, NT);
    if (vec) {
        auto k = sos_warm_kernel<NSEC, REV, true, NT, NUM>;
        ECOG_TRY((smem_attr<sos_warm_kernel<NSEC, REV, true, NT, NUM>>(smem)));
        k<<<grid, NT, smem, st>>>(x, y, C, T, ldx, ldy, p.chunk, p.tail, nChunks, p.padlen, p.zero_phase, coef, padbuf, gain, 0);
    } else {
        auto k = sos_warm_kernel<NSEC, REV, false, NT, NUM>;
        ECOG_TRY((smem_attr<sos_warm_kernel<NSEC, REV, false, NT, NUM>>(smem)));
        k<<<grid, NT, smem, st>>>(x, y, C, T, ldx, ldy, p.chunk, p.tail, nChunks, p.padlen, p.zero_phase, coef, padbuf, gain, 0);
    }
    return check_launch(REV ? "sos_warm_bwd" : "sos_warm_fwd");
}

// cascade pair (two 4-section unit-form cascades as one sweep): 128-bit copies only
template <int NSEC>
static int run_sos_warm(const float* x, float* y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                        const ecog_sos_plan& p, const SosCoef& coef_in, float* tmp, int64_t ldt, double* padbuf,
                        cudaStream_t st) {
    const int nChunks = (int)ceil_div(T, p.chunk);
    float* mid = p.zero_phase ? tmp : y;
    const int64_t ldm = p.zero_phase ? ldt : ldy;
    const bool vec = aligned16(x) && aligned16(y) && aligned16(mid) && T % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0 && ldm % 4 == 0;
    const bool wide = p.threads >= 512;
    // numerator form: unit sections g (1 + beta1 z^-1 +- z^-2) with the gain on section 0 only
    SosCoef coef = coef_in;
    int num = unit_form(coef, 0, NSEC, 0);
    if (num == 8 && NSEC != 4) num = 0;          // the direct-form variant is instantiated for 4 sections
    const double gain = prepare_form(coef, 0, NSEC, 0, num);
#define ECOG_WARM(REVV, NTT, Y_IN, Y_OUT, LD_IN, LD_OUT)                                                                         \
    (num == 2 ? launch_warm<NSEC, REVV, NTT, 2>(Y_IN, Y_OUT, C, T, LD_IN, LD_OUT, p, nChunks, coef, gain, padbuf, vec, st)       \
     : num == 5 ? launch_warm<NSEC, REVV, NTT, 5>(Y_IN, Y_OUT, C, T, LD_IN, LD_OUT, p, nChunks, coef, gain, padbuf, vec, st)     \
     : num == 8 ? launch_warm<NSEC, REVV, NTT, (NSEC == 4 ? 8 : 0)>(Y_IN, Y_OUT, C, T, LD_IN, LD_OUT, p, nChunks, coef, gain, padbuf, vec, st) \
                : launch_warm<NSEC, REVV, NTT, 0>(Y_IN, Y_OUT, C, T, LD_IN, LD_OUT, p, nChunks, coef, gain, padbuf, vec, st))
    if (wide) ECOG_TRY(ECOG_WARM(false, 512, x, mid, ldx, ldm));
    else      ECOG_TRY(ECOG_WARM(false, 256, x, mid, ldx, ldm));
    if (!p.zero_phase) return ECOG_OK;
    if (wide) return ECOG_WARM(true, 512, mid, y, ldm, ldy);
    return ECOG_WARM(true, 256, mid, y, ldm, ldy);
#undef ECOG_WARM
}

static size_t sos_smem_bytes() {
    return ((size_t)kRing * kSosThreads * kPitch) * sizeof(float) +
           (size_t)kSosThreads * (sizeof(int64_t) + 2 * sizeof(int));
}

template <int NSEC, bool REV, bool WRITE>
static int launch_chunk(const float* x, float* y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                        const ecog_sos_plan& p, int nChunks, const SosCoef& coef, double* state,
                        double* gbuf, double* padbuf, bool vec, cudaStream_t st) {
    const size_t smem = sos_smem_bytes();
    const unsigned grid = (unsigned)ceil_div(C * nChunks, kSosThreads);
    if (vec) {
        auto k = sos_chunk_kernel<NSEC, REV, WRITE, true>;
        ECOG_TRY((smem_attr<sos_chunk_kernel<NSEC, REV, WRITE, true>>(smem)));
        k<<<grid, kSosThreads, smem, st>>>(x, y, C, T, ldx, ldy, p.chunk, p.tail, nChunks, p.padlen, coef, state, gbuf, padbuf);
    } else {
        auto k = sos_chunk_kernel<NSEC, REV, WRITE, false>;
        ECOG_TRY((smem_attr<sos_chunk_kernel<NSEC, REV, WRITE, false>>(smem)));
        k<<<grid, kSosThreads, smem, st>>>(x, y, C, T, ldx, ldy, p.chunk, p.tail, nChunks, p.padlen, coef, state, gbuf, padbuf);
    }
    return check_launch(WRITE ? "sos_main" : "sos_tail");
}

template <int NSEC>
static int run_sos(const float* x, float* y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                   const ecog_sos_plan& p, const SosCoef& coef, const SosMatrix& M, double* state,
                   double* gbuf, double* padbuf, cudaStream_t st) {
    const int nChunks = (int)ceil_div(T, p.chunk);
    const bool vec = aligned16(x) && aligned16(y) && T % 4 == 0 && ldx % 4 == 0 && ldy % 4 == 0;
    const unsigned sgrid = (unsigned)ceil_div(C, kScanWarps);
    // forward sweep: x -> y
    if (nChunks > 1) ECOG_TRY((launch_chunk<NSEC, false, false>(x, y, C, T, ldx, ldy, p, nChunks, coef, state, gbuf, padbuf, vec, st)));
    sos_scan_kernel<NSEC, false><<<sgrid, kScanWarps * 32, 0, st>>>(x, C, T, ldx, nChunks, p.padlen, p.zero_phase, coef, M, state, gbuf, padbuf);
    ECOG_TRY(check_launch("sos_scan"));
    ECOG_TRY((launch_chunk<NSEC, false, true>(x, y, C, T, ldx, ldy, p, nChunks, coef, state, gbuf, padbuf, vec, st)));
    if (!p.zero_phase) return ECOG_OK;
    // backward sweep: y -> y in place, time reversed
    if (nChunks > 1) ECOG_TRY((launch_chunk<NSEC, true, false>(y, y, C, T, ldy, ldy, p, nChunks, coef, state, gbuf, padbuf, vec, st)));
    sos_scan_kernel<NSEC, true><<<sgrid, kScanWarps * 32, 0, st>>>(y, C, T, ldy, nChunks, p.padlen, p.zero_phase, coef, M, state, gbuf, padbuf);
    ECOG_TRY(check_launch("sos_scan"));
    return launch_chunk<NSEC, true, true>(y, y, C, T, ldy, ldy, p, nChunks, coef, state, gbuf, padbuf, vec, st);
}

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

}  // namespace ecog

using namespace ecog;

static int64_t warm_ld(int64_t T) { return (T + 3) / 4 * 4; }

extern "C" size_t ecog_sos_workspace(const ecog_sos_plan* plan, int64_t C, int64_t T) {
    if (!plan || plan->chunk <= 0) return 0;
    if (plan->mode == ECOG_SOS_WARMUP || plan->mode == ECOG_SOS_WARMUP_TMA) {      // filtered right pad + (zero phase) the forward result
        size_t pad = align_up((size_t)C * (plan->padlen > 0 ? plan->padlen : 1) * sizeof(double), 256);
        return pad + (plan->zero_phase ? (size_t)C * warm_ld(T) * sizeof(float) : 0);
    }
    const int64_t nChunks = ceil_div(T, plan->chunk);
    const int nsec_pad = plan->nsec <= 4 ? (plan->nsec == 3 ? 3 : (plan->nsec <= 2 ? plan->nsec : 4))
                                         : (plan->nsec <= 6 ? 6 : 8);
    size_t states = align_up((size_t)C * nChunks * 2 * nsec_pad * sizeof(double), 256);
    size_t pad = align_up((size_t)C * (plan->padlen > 0 ? plan->padlen : 1) * sizeof(double), 256);
    return 2 * states + pad;       // chunk start states, chunk affine terms, filtered right pad
}

extern "C" int ecog_sosfilt(const float* d_x, float* d_y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                            const ecog_sos_plan* plan, const double* h_sos, const double* h_zi,
                            const double* h_M, void* d_workspace, size_t workspace_bytes,
                            ecog_stream_t stream) {
    if (!plan || !h_sos) return fail(ECOG_E_VALUE, "ecog_sosfilt: null plan");
    const ecog_sos_plan& p = *plan;
    if (C <= 0 || T <= 0 || ldx < T || ldy < T || T >= (int64_t)1 << 31)
        return fail(ECOG_E_VALUE, "ecog_sosfilt: bad shape C=%lld T=%lld", (long long)C, (long long)T);
    if (p.nsec < 1 || p.nsec > ECOG_MAX_SECTIONS) return fail(ECOG_E_VALUE, "ecog_sosfilt: nsec=%d out of range", p.nsec);
    if (p.chunk < kSub || p.chunk % kSub || p.tail < 0 || p.tail % kSub)
        return fail(ECOG_E_VALUE, "ecog_sosfilt: chunk=%d tail=%d must be multiples of %d", p.chunk, p.tail, kSub);
    if (p.mode == ECOG_SOS_SCAN && (p.tail < kSub || p.tail > p.chunk))
        return fail(ECOG_E_VALUE, "ecog_sosfilt: scan mode needs %d <= tail=%d <= chunk=%d", kSub, p.tail, p.chunk);
    if (p.zero_phase) {
        if (!h_zi) return fail(ECOG_E_VALUE, "ecog_sosfilt: zero_phase needs h_zi");
        if (p.padlen < 1) return fail(ECOG_E_VALUE, "ecog_sosfilt: zero_phase needs padlen >= 1");
        if (T <= p.padlen)
            return fail(ECOG_E_VALUE, "The length of the input vector x must be greater than padlen, which is %d.", p.padlen);
    }
    const int nChunks = (int)ceil_div(T, p.chunk);
    if (p.mode != ECOG_SOS_SCAN && p.mode != ECOG_SOS_WARMUP && p.mode != ECOG_SOS_WARMUP_TMA)
        return fail(ECOG_E_VALUE, "ecog_sosfilt: unknown mode %d", p.mode);
    if (p.split && p.mode == ECOG_SOS_SCAN) return fail(ECOG_E_VALUE, "ecog_sosfilt: the cascade pair runs in warm-up mode only");
    if (p.mode == ECOG_SOS_WARMUP_TMA) {
        if (!p.zero_phase || p.chunk % 32 || p.tail % 32 || p.tail > p.chunk || T % p.chunk || ldx != T || ldy != T ||
            !aligned16(d_x) || !aligned16(d_y) || !aligned16(d_workspace))
            return fail(ECOG_E_UNSUPPORTED, "ecog_sosfilt (TMA): needs zero phase, contiguous 16-byte aligned rows, "
                                           "T a multiple of chunk, chunk and tail multiples of 32, tail <= chunk");
    }
    if (p.mode == ECOG_SOS_WARMUP && !p.zero_phase && d_x == d_y)
        return fail(ECOG_E_VALUE, "ecog_sosfilt: the causal warm-up path cannot run in place");
    if (p.mode == ECOG_SOS_SCAN && nChunks > 1 && !h_M)
        return fail(ECOG_E_VALUE, "ecog_sosfilt: chunked rows need the chunk transition matrix h_M");
    if (workspace_bytes < ecog_sos_workspace(plan, C, T))
        return fail(ECOG_E_WORKSPACE, "ecog_sosfilt: workspace %zu < %zu", workspace_bytes, ecog_sos_workspace(plan, C, T));

    // pad the cascade with identity sections up to an instantiated width
    const int widths[] = {1, 2, 3, 4, 6, 8};
    int ns = 8;
    for (int w : widths) if (p.nsec <= w) { ns = w; break; }
    SosCoef coef;
    SosMatrix M;
    memset(&coef, 0, sizeof(coef));
    memset(&M, 0, sizeof(M));
    for (int j = 0; j < ns; ++j) {
        if (j < p.nsec) {
            const double* r = h_sos + 6 * j;
            if (r[3] == 0.0) return fail(ECOG_E_VALUE, "ecog_sosfilt: a0 of section %d is zero", j);
            const double a0 = r[3];
            coef.c[j][0] = r[0] / a0; coef.c[j][1] = r[1] / a0; coef.c[j][2] = r[2] / a0;
            coef.c[j][3] = r[4] / a0; coef.c[j][4] = r[5] / a0;
            if (h_zi) { coef.zi[j][0] = h_zi[2 * j]; coef.zi[j][1] = h_zi[2 * j + 1]; }
        } else {
            coef.c[j][0] = 1.0;   // identity section: y = u, state stays zero
        }
    }
    if (h_M) {
        const int n0 = 2 * p.nsec;
        for (int i = 0; i < n0; ++i)
            for (int j = 0; j < n0; ++j) M.m[i][j] = h_M[i * n0 + j];
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (p.mode == ECOG_SOS_WARMUP_TMA) {
        double* padbuf = (double*)d_workspace;
        const size_t pad = align_up((size_t)C * (p.padlen > 0 ? p.padlen : 1) * sizeof(double), 256);
        if (p.split && ((p.split & ~ECOG_SOS_SPLIT_F32B) != 4 || p.nsec != 8 || p.tail_b < 0 || p.tail_b % 32 || p.tail_b > p.tail))
            return fail(ECOG_E_VALUE, "ecog_sosfilt (TMA): cascade pair needs nsec=8, split=4, tail_b a multiple of 32 <= tail");
        return run_sos_warm_tma(d_x, d_y, C, T, p, coef, (float*)((char*)d_workspace + pad), padbuf, st);
    }
    if (p.mode == ECOG_SOS_WARMUP) {
        double* padbuf = (double*)d_workspace;
        const size_t pad = align_up((size_t)C * (p.padlen > 0 ? p.padlen : 1) * sizeof(double), 256);
        float* tmp = (float*)((char*)d_workspace + pad);
        const int64_t ldt = warm_ld(T);
        if (p.split) {
            if (p.split != 4 || p.nsec != 8 || p.tail_b < 0 || p.tail_b % kSub || p.tail_b > p.tail)
                return fail(ECOG_E_VALUE, "ecog_sosfilt: cascade pair needs nsec=8, split=4 (the float32 half exists on the TMA path only), "
                                          "0 <= tail_b <= tail (multiples of %d)", kSub);
            return run_sos_warm_pair(d_x, d_y, C, T, ldx, ldy, p, coef, tmp, ldt, padbuf, st);
        }
        switch (ns) {
            case 1: return run_sos_warm<1>(d_x, d_y, C, T, ldx, ldy, p, coef, tmp, ldt, padbuf, st);
            case 2: return run_sos_warm<2>(d_x, d_y, C, T, ldx, ldy, p, coef, tmp, ldt, padbuf, st);
            case 3: return run_sos_warm<3>(d_x, d_y, C, T, ldx, ldy, p, coef, tmp, ldt, padbuf, st);
            case 4: return run_sos_warm<4>(d_x, d_y, C, T, ldx, ldy, p, coef, tmp, ldt, padbuf, st);
            case 6: return run_sos_warm<6>(d_x, d_y, C, T, ldx, ldy, p, coef, tmp, ldt, padbuf, st);
            default: return run_sos_warm<8>(d_x, d_y, C, T, ldx, ldy, p, coef, tmp, ldt, padbuf, st);
        }
    }
    const size_t states = align_up((size_t)C * nChunks * 2 * ns * sizeof(double), 256);
    double* state = (double*)d_workspace;
    double* gbuf = (double*)((char*)d_workspace + states);
    double* padbuf = (double*)((char*)d_workspace + 2 * states);
    switch (ns) {
        case 1: return run_sos<1>(d_x, d_y, C, T, ldx, ldy, p, coef, M, state, gbuf, padbuf, st);
        case 2: return run_sos<2>(d_x, d_y, C, T, ldx, ldy, p, coef, M, state, gbuf, padbuf, st);
        case 3: return run_sos<3>(d_x, d_y, C, T, ldx, ldy, p, coef, M, state, gbuf, padbuf, st);
        case 4: return run_sos<4>(d_x, d_y, C, T, ldx, ldy, p, coef, M, state, gbuf, padbuf, st);
        case 6: return run_sos<6>(d_x, d_y, C, T, ldx, ldy, p, coef, M, state, gbuf, padbuf, st);
        default: return run_sos<8>(d_x, d_y, C, T, ldx, ldy, p, coef, M, state, gbuf, padbuf, st);
    }
}
