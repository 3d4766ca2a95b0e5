// K5a circular FIR low-pass + integer decimation: first stage of the two-stage resampler.
//
//   y[c, m] = sum_{j < ntaps} h[j] * x[c, (m*D + j - off) mod T],   m = 0 .. T/D - 1
//
// scipy.signal.resample (ref: preprocess/signal/downsample.py:21-27) is a brick-wall in the
// frequency domain of the WHOLE row.  Every bin the brick wall keeps (k <= num/2) survives this
// stage unaliased to 120 dB (Kaiser stop band from T/D - num/2 on), and its known gain H1[k] is
// divided out exactly by the FFT stage that follows (ecog_resample_tables::bin_gain), so the
// pair reproduces the brick wall while the global FFT only sees T/D samples per row.
//
// One CTA = one channel x 256*R consecutive outputs.  The input span is staged in shared memory
// as float4 words with one pad word every S4 = D*R/4 words, so thread t's window (which starts
// S4 words after thread t-1's) is read with conflict-free 128-bit loads.  Every thread slides
// its window once and feeds R accumulators; the taps are kernel parameters (constant bank), the
// tap loop is fully unrolled, so the inner loop is pure FFMA with constant operands:
// ntaps/D FMA per input sample (40 for 2 kHz -> 500 Hz), HBM traffic 4 + 4/D bytes per sample.
#include "common.cuh"

namespace ecog {

constexpr int kFirThreads = 256;
constexpr int kFirMaxTaps = 256;

struct FirTaps { float h[kFirMaxTaps]; };

template <int D, int NT4, int R, bool VEC>
__global__ void __launch_bounds__(kFirThreads)
fir_decimate_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t T, int64_t T1,
                    int64_t ldx, int64_t ldy, int off, const __grid_constant__ FirTaps taps) {
    constexpr int NTAPS = 4 * NT4;
    constexpr int S4 = D * R / 4;                                 // window stride between threads, float4 words
    static_assert((D * R) % 4 == 0 && S4 % 2 == 0, "thread stride must be an even number of float4 words");
    constexpr int W4 = (D * (R - 1) + NTAPS + 3) / 4;             // words one thread reads
    constexpr int SPAN4 = S4 * (kFirThreads - 1) + W4;            // words the CTA needs
    extern __shared__ __align__(16) float4 xs[];                 // [SPAN4 + SPAN4 / S4 + 1]
    const int tid = threadIdx.x;
    const int64_t ch = blockIdx.y;
    const int64_t m0 = (int64_t)blockIdx.x * (kFirThreads * R);   // first output of this CTA
    const float* xr = x + ch * ldx;
    const int64_t tin = m0 * D - off;                             // first input sample staged (may be < 0)

    const bool interior = tin >= 0 && tin + 4 * (int64_t)SPAN4 <= T;
    if (VEC && interior) {
        for (int n = tid; n < SPAN4; n += kFirThreads) cp_async16(&xs[n + n / S4], xr + tin + 4 * (int64_t)n);
        cp_async_commit();
        cp_async_wait<0>();
    } else {
        for (int n = tid; n < SPAN4; n += kFirThreads) {
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                int64_t t = (tin + 4 * (int64_t)n + e) % T;       // circular, like the whole-row FFT
                if (t < 0) t += T;
                v[e] = xr[t];
            }
            xs[n + n / S4] = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
    __syncthreads();

    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;
    const float4* win = xs + (S4 + 1) * tid;                      // pad(S4*tid + w) = (S4+1)*tid + w + w/S4
#pragma unroll
    for (int w = 0; w < W4; ++w) {
        const float4 v4 = win[w + w / S4];
        const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int j = 4 * w + e - D * r;                  // compile-time tap index
                if (j >= 0 && j < NTAPS) acc[r] = fmaf(taps.h[j], v[e], acc[r]);
            }
        }
    }
    const int64_t m = m0 + (int64_t)R * tid;
    float* yr = y + ch * ldy + m;
    if (VEC && m + R <= T1) {
#pragma unroll
        for (int r = 0; r < R; r += 4)
            *reinterpret_cast<float4*>(yr + r) = make_float4(acc[r], acc[r + 1], acc[r + 2], acc[r + 3]);
    } else {
#pragma unroll
        for (int r = 0; r < R; ++r)
            if (m + r < T1) yr[r] = acc[r];
    }
}

template <int D, int NT4>
static int launch_fir(const float* x, float* y, int64_t C, int64_t T, int64_t ldx, int64_t ldy, int off,
                      const FirTaps& taps, bool vec, cudaStream_t st) {
    constexpr int R = 8;
    constexpr int S4 = D * R / 4;
    constexpr int W4 = (D * (R - 1) + 4 * NT4 + 3) / 4;
    constexpr int SPAN4 = S4 * (kFirThreads - 1) + W4;
    const size_t smem = (size_t)(SPAN4 + SPAN4 / S4 + 1) * sizeof(float4);
    const int64_t T1 = T / D;
    dim3 grid((unsigned)ceil_div(T1, (int64_t)kFirThreads * R), (unsigned)C);
    if (vec) {
        auto k = fir_decimate_kernel<D, NT4, R, true>;
        ECOG_TRY((smem_attr<fir_decimate_kernel<D, NT4, R, true>>(smem)));
        k<<<grid, kFirThreads, smem, st>>>(x, y, T, T1, ldx, ldy, off, taps);
    } else {
        auto k = fir_decimate_kernel<D, NT4, R, false>;
        ECOG_TRY((smem_attr<fir_decimate_kernel<D, NT4, R, false>>(smem)));
        k<<<grid, kFirThreads, smem, st>>>(x, y, T, T1, ldx, ldy, off, taps);
    }
    return check_launch("fir_decimate");
}

template <int D>
static int dispatch_taps(int nt4, const float* x, float* y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                         int off, const FirTaps& taps, bool vec, cudaStream_t st) {
    if (nt4 <= 8) return launch_fir<D, 8>(x, y, C, T, ldx, ldy, off, taps, vec, st);
    if (nt4 <= 16) return launch_fir<D, 16>(x, y, C, T, ldx, ldy, off, taps, vec, st);
    if (nt4 <= 24) return launch_fir<D, 24>(x, y, C, T, ldx, ldy, off, taps, vec, st);
    if (nt4 <= 32) return launch_fir<D, 32>(x, y, C, T, ldx, ldy, off, taps, vec, st);
    if (nt4 <= 40) return launch_fir<D, 40>(x, y, C, T, ldx, ldy, off, taps, vec, st);
    if (nt4 <= 48) return launch_fir<D, 48>(x, y, C, T, ldx, ldy, off, taps, vec, st);
    return launch_fir<D, 64>(x, y, C, T, ldx, ldy, off, taps, vec, st);
}

// ------------------------------------------------------------------ K6 causal FIR (long)
// replaces the `fir` band method (ref: preprocess/signal/frequency_filter.py:260-274): the mean
// over centre frequencies of lfilter(firwin(order+1, ...), 1, x) is ONE causal FIR with the
// averaged taps (linearity), zero initial state:
//   y[t] = sum_{j <= order} h[j] x[t - j],  x[t < 0] = 0.
// Run as a correlation with the reversed taps g[i] = h[n-1-i] over a staged tile; taps live in
// shared memory (read as broadcast 128-bit loads), every thread slides a 12-sample register
// window over its 8 outputs: 32 FMA per two shared loads.
constexpr int kFirR = 8;

__global__ void __launch_bounds__(kFirThreads)
fir_causal_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t T, int64_t ldx, int64_t ldy,
                  const float* __restrict__ g, int n4, bool vec) {
    // g: 4*n4 reversed taps (zero padded at the end), output t uses x[t - off + i], off = 4*n4 - 4:
    // the host pads h so that the last real tap sits at index off (a multiple of 4).
    extern __shared__ __align__(16) float4 fsm[];
    constexpr int S4 = kFirR / 4;                                  // 2 words between neighbouring threads
    const int span4 = S4 * kFirThreads + n4 + 1;                   // words staged per CTA
    float4* xs = fsm;                                              // [span4 + span4 / S4 + 1], pad word every S4
    float4* gs = fsm + span4 + span4 / S4 + 1;                     // [n4]
    const int tid = threadIdx.x;
    const int64_t ch = blockIdx.y;
    const int64_t m0 = (int64_t)blockIdx.x * (kFirThreads * kFirR);
    const float* xr = x + ch * ldx;
    const int64_t tin = m0 - (int64_t)(4 * n4 - 4);                // first staged sample (may be < 0: zeros)
    for (int n = tid; n < n4; n += kFirThreads) gs[n] = *reinterpret_cast<const float4*>(g + 4 * n);
    for (int n = tid; n < span4; n += kFirThreads) {
        const int64_t t = tin + 4 * (int64_t)n;
        float4 v;
        if (vec && t >= 0 && t + 4 <= T) {
            v = *reinterpret_cast<const float4*>(xr + t);
        } else {
            float e[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) e[k] = (t + k >= 0 && t + k < T) ? xr[t + k] : 0.f;
            v = make_float4(e[0], e[1], e[2], e[3]);
        }
        xs[n + n / S4] = v;
    }
    __syncthreads();
    float acc[kFirR];
#pragma unroll
    for (int r = 0; r < kFirR; ++r) acc[r] = 0.f;
    const float4* win = xs + (S4 + 1) * tid;                       // word w of this thread: win[w + w / S4]
    float4 a = win[0], b = win[1 + 1 / S4];                        // thread-relative words 0 and 1
    for (int w = 0; w < n4; ++w) {
        const float4 c = win[(w + 2) + (w + 2) / S4];              // word w + 2: samples 4w+8 .. 4w+11
        const float4 gw = gs[w];
        const float xv[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
        const float gv[4] = {gw.x, gw.y, gw.z, gw.w};
#pragma unroll
        for (int r = 0; r < kFirR; ++r)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[r] = fmaf(gv[e], xv[r + e], acc[r]);
        a = b;
        b = c;
    }
    const int64_t m = m0 + (int64_t)kFirR * tid;
    float* yr = y + ch * ldy + m;
    if (vec && m + kFirR <= T) {
        *reinterpret_cast<float4*>(yr) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        *reinterpret_cast<float4*>(yr + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    } else {
#pragma unroll
        for (int r = 0; r < kFirR; ++r)
            if (m + r < T) yr[r] = acc[r];
    }
}

}  // namespace ecog

using namespace ecog;

extern "C" int ecog_fir_decimate(const float* d_x, float* d_y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                                 const float* h_taps, int32_t ntaps, int32_t offset, int32_t D,
                                 ecog_stream_t stream) {
    if (C <= 0 || C > 65535 || T <= 0 || ldx < T) return fail(ECOG_E_VALUE, "ecog_fir_decimate: bad shape");
    if (D != 2 && D != 4) return fail(ECOG_E_UNSUPPORTED, "ecog_fir_decimate: decimation factor %d not built (2, 4)", D);
    if (T % D) return fail(ECOG_E_VALUE, "ecog_fir_decimate: T=%lld is not a multiple of D=%d", (long long)T, D);
    if (ldy < T / D) return fail(ECOG_E_VALUE, "ecog_fir_decimate: output stride too small");
    if (!h_taps || ntaps < 1 || ntaps > kFirMaxTaps)
        return fail(ECOG_E_VALUE, "ecog_fir_decimate: 1..%d taps supported, got %d", kFirMaxTaps, ntaps);
    if (offset < 0 || offset >= T) return fail(ECOG_E_VALUE, "ecog_fir_decimate: bad tap offset %d", offset);
    if (d_x == d_y) return fail(ECOG_E_VALUE, "ecog_fir_decimate: in-place operation is not supported");
    FirTaps taps;
    memset(&taps, 0, sizeof(taps));
    for (int j = 0; j < ntaps; ++j) taps.h[j] = h_taps[j];
    const int nt4 = (ntaps + 3) / 4;
    const bool vec = aligned16(d_x) && aligned16(d_y) && ldx % 4 == 0 && ldy % 4 == 0 && offset % 4 == 0;
    cudaStream_t st = (cudaStream_t)stream;
    if (D == 2) return dispatch_taps<2>(nt4, d_x, d_y, C, T, ldx, ldy, offset, taps, vec, st);
    return dispatch_taps<4>(nt4, d_x, d_y, C, T, ldx, ldy, offset, taps, vec, st);
}

extern "C" int ecog_fir_causal(const float* d_x, float* d_y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                               const float* d_taps_rev, int32_t ntaps4, ecog_stream_t stream) {
    if (C <= 0 || C > 65535 || T <= 0 || ldx < T || ldy < T) return fail(ECOG_E_VALUE, "ecog_fir_causal: bad shape");
    if (!d_taps_rev || ntaps4 < 1 || ntaps4 > 2048)
        return fail(ECOG_E_VALUE, "ecog_fir_causal: 1..2048 tap words supported, got %d", ntaps4);
    if (!aligned16(d_taps_rev)) return fail(ECOG_E_VALUE, "ecog_fir_causal: tap table must be 16-byte aligned");
    if (d_x == d_y) return fail(ECOG_E_VALUE, "ecog_fir_causal: in-place operation is not supported");
    const bool vec = aligned16(d_x) && aligned16(d_y) && ldx % 4 == 0 && ldy % 4 == 0;
    const int span4 = (kFirR / 4) * kFirThreads + ntaps4 + 1;
    const size_t smem = (size_t)(span4 + span4 / (kFirR / 4) + 1 + ntaps4) * sizeof(float4);
    ECOG_TRY((smem_attr<fir_causal_kernel>(smem)));
    dim3 grid((unsigned)ceil_div(T, (int64_t)kFirThreads * kFirR), (unsigned)C);
    fir_causal_kernel<<<grid, kFirThreads, smem, (cudaStream_t)stream>>>(d_x, d_y, T, ldx, ldy, d_taps_rev, ntaps4, vec);
    return check_launch("fir_causal");
}
