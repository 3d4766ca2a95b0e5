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

// ------------------------------------------------------------------ K5a', decimation by 4 as TWO half-band stages
// The band edges of a decimate-by-2 stage in front of a brick wall at f_pass are f_pass and 1 - f_pass (what folds
// onto the kept band), symmetric about half the Nyquist frequency: BOTH stages are half-band filters -- every second
// tap is zero except the centre one:
//     y[n] = hc x[2n] + sum_{i < K} g_i (x[2n - 2i - 1] + x[2n + 2i + 1])              (2K + 1 products per output)
// Stage 1 (T -> T/2: transition f_pass .. 1 - f_pass, wide) needs K1 = 8, stage 2 (T/2 -> T/4: 2 f_pass .. 1 - 2 f_pass)
// K2 = 21 for 120 dB at f_pass = 0.2 (2 kHz -> 400 Hz): 17/2 + 43/4 = 19.25 products per input sample against the
// 40 of the single 160-tap stage, which takes the step off the FP32 pipe.  The intermediate row lives
// in shared memory only.  Same circular indexing and the same exact compensation as above: the FFT stage divides the
// kept bins by H1[k] H2[k] (fftplan.predecimation).
// One CTA = one channel x 960 outputs (120 threads x 8): 2016 stage-1 values (126 tasks of 16, one per thread -- a
// second round for a few stragglers would idle the CTA at the barrier) from 4064 staged inputs; both windows are
// read with conflict-free 128-bit loads (one pad word per thread stride, as above).
constexpr int kHbK1 = 8, kHbK2 = 21;
struct HbTaps { float c1, g1[kHbK1], c2, g2[kHbK2]; };

// 128-bit shared load the compiler may not narrow: a half-band stage uses only every second float of most words, and
// scalar loads at a lane stride of 80 / 144 bytes conflict four ways (ncu: 3.9 wavefronts per LDS, 56 % of the
// kernel's shared wavefronts were conflicts) where the padded 128-bit pattern has none
__device__ __forceinline__ float4 lds128(const float4* p) {
    float4 v;
    asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];"      // .volatile: ptxas narrows a plain v4 load whose components are dead
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
}

template <int K, int R, int C0, int NW, int S4>
__device__ __forceinline__ void hb_accumulate(const float4* __restrict__ win, float hc, const float (&g)[K], float (&acc)[R]) {
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        const float4 v4 = lds128(win + w + w / S4);
        const float v[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const int d = 4 * w + e - (C0 + 2 * r);           // compile-time distance from output r's centre
                if (d == 0) acc[r] = fmaf(hc, v[e], acc[r]);
                else if ((d & 1) && d < 2 * K && d > -2 * K) acc[r] = fmaf(g[((d < 0 ? -d : d) - 1) / 2], v[e], acc[r]);
            }
        }
    }
}

constexpr int kHbR2 = 8;                                   // outputs per thread
constexpr int kHbThreads = 128;                            // threads per CTA: small CTAs, seven per SM -- the kernel's three
                                                           // phases are separated by CTA barriers, and what hides them is other CTAs
                                                           // (256 threads x 4 CTAs: 2.48 ms, 128 x 7: 2.35 ms, 64 x 16: 2.36 ms at C2;
                                                           // double-buffered inputs at 256 x 2: 2.58 ms)
constexpr int kHbT2 = kHbThreads - 8;                      // threads with stage-2 work: their stage-1 values are <= kHbThreads tasks of 16
constexpr int kHbM2 = kHbT2 * kHbR2;                       // 960 outputs per tile
constexpr int kHbL1 = (2 * kHbK1 - 1 + 3) / 4 * 4;         // window lead-ins: multiples of 4 >= 2K - 1 (16, 44)
constexpr int kHbL2 = (2 * kHbK2 - 1 + 3) / 4 * 4;
constexpr int kHbNW1 = (kHbL1 + 2 * 15 + 2 * kHbK1 - 1) / 4 + 1;               // window words of a stage-1 task (16)
constexpr int kHbNW2 = (kHbL2 + 2 * (kHbR2 - 1) + 2 * kHbK2 - 1) / 4 + 1;      // window words of a stage-2 thread (25)
constexpr int kHbTasks1 = (16 * (kHbT2 - 1) + 4 * kHbNW2 + 15) / 16;           // stage-1 tasks of 16 values (126)
static_assert(kHbTasks1 <= kHbThreads, "one stage-1 task per thread: a second round would idle the CTA at the barrier");
constexpr int kHbXW = 8 * (kHbTasks1 - 1) + kHbNW1;        // input words (float4) per tile
constexpr int kHbXS = kHbXW + kHbXW / 8 + 1;               // padded
constexpr int kHbYW = 4 * kHbTasks1;                       // stage-1 words
constexpr int kHbYS = kHbYW + kHbYW / 4 + 1;
constexpr int kHbTilesPerCta = 8;

__global__ void __launch_bounds__(kHbThreads, 7)
halfband2_decimate_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t T, int64_t T4, int64_t ldx,
                          int64_t ldy, bool vec, const __grid_constant__ HbTaps taps) {
    extern __shared__ __align__(16) float4 hsm[];
    float4* xs = hsm;                                      // [kHbXS]  inputs, one pad word every 8
    float4* ys = hsm + kHbXS;                              // [kHbYS]  stage-1 values, one pad word every 4
    const int tid = threadIdx.x;
    const int64_t ch = blockIdx.y;
    const float* xr = x + ch * ldx;
    const int64_t nTiles = (T4 + kHbM2 - 1) / kHbM2;
    // a CTA walks kHbTilesPerCta consecutive tiles of its row (fewer, longer-lived CTAs: 2.35 against 2.41 ms at C2)
    for (int64_t tile = (int64_t)blockIdx.x * kHbTilesPerCta; tile < nTiles && tile < (int64_t)(blockIdx.x + 1) * kHbTilesPerCta; ++tile) {
    const int64_t m0 = tile * kHbM2;                       // first output of this tile
    const int64_t tin = 2 * (2 * m0 - kHbL2) - kHbL1;      // first input sample staged (may be < 0), multiple of 4

    // ---- inputs: asynchronous 16-byte copies inside the row, scalar circular loads for the first / last tiles
    if (vec && tin >= 0 && tin + 4 * (int64_t)kHbXW <= T) {
        for (int n = tid; n < kHbXW; n += kHbThreads) cp_async16(&xs[n + n / 8], xr + tin + 4 * (int64_t)n);
        cp_async_commit();
        cp_async_wait<0>();
    } else {
        for (int n = tid; n < kHbXW; n += kHbThreads) {
            float v[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                int64_t t = (tin + 4 * (int64_t)n + e) % T;        // circular, like the whole-row FFT
                if (t < 0) t += T;
                v[e] = xr[t];
            }
            xs[n + n / 8] = make_float4(v[0], v[1], v[2], v[3]);
        }
    }
    __syncthreads();

    // ---- stage 1: task k -> stage-1 values 16 k .. 16 k + 15 (window: input words 8 k .. 8 k + 15, centre of value 0 at float kHbL1)
    if (const int k = tid; k < kHbTasks1) {
        float acc[16];
#pragma unroll
        for (int r = 0; r < 16; ++r) acc[r] = 0.f;
        hb_accumulate<kHbK1, 16, kHbL1, kHbNW1, 8>(xs + 9 * k, taps.c1, taps.g1, acc);
#pragma unroll
        for (int j = 0; j < 4; ++j) ys[5 * k + j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
    }
    __syncthreads();

    // ---- stage 2: thread t -> outputs 8 t .. 8 t + 7 (window: stage-1 words 4 t .. 4 t + 24, centre of output 0 at float kHbL2)
    if (tid < kHbT2) {
        float acc[kHbR2];
#pragma unroll
        for (int r = 0; r < kHbR2; ++r) acc[r] = 0.f;
        hb_accumulate<kHbK2, kHbR2, kHbL2, kHbNW2, 4>(ys + 5 * tid, taps.c2, taps.g2, acc);
        const int64_t m = m0 + (int64_t)kHbR2 * tid;
        float* yr = y + ch * ldy + m;
        if (vec && m + kHbR2 <= T4) {
            *reinterpret_cast<float4*>(yr) = make_float4(acc[0], acc[1], acc[2], acc[3]);
            *reinterpret_cast<float4*>(yr + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
        } else {
#pragma unroll
            for (int r = 0; r < kHbR2; ++r)
                if (m + r < T4) yr[r] = acc[r];
        }
    }
    __syncthreads();            // this tile's reads of xs / ys are done before the next tile refills them
    }
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

extern "C" int ecog_halfband2_decimate(const float* d_x, float* d_y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                                       const float* h_stage1, int32_t k1, const float* h_stage2, int32_t k2,
                                       ecog_stream_t stream) {
    if (C <= 0 || C > 65535 || T <= 0 || ldx < T) return fail(ECOG_E_VALUE, "ecog_halfband2_decimate: bad shape");
    if (T % 4) return fail(ECOG_E_VALUE, "ecog_halfband2_decimate: T=%lld is not a multiple of 4", (long long)T);
    if (ldy < T / 4) return fail(ECOG_E_VALUE, "ecog_halfband2_decimate: output stride too small");
    if (!h_stage1 || !h_stage2 || k1 < 1 || k1 > kHbK1 || k2 < 1 || k2 > kHbK2)
        return fail(ECOG_E_UNSUPPORTED, "ecog_halfband2_decimate: stages of up to %d and %d odd tap pairs are built (got %d, %d)",
                    kHbK1, kHbK2, k1, k2);
    if (d_x == d_y) return fail(ECOG_E_VALUE, "ecog_halfband2_decimate: in-place operation is not supported");
    HbTaps taps;
    memset(&taps, 0, sizeof(taps));
    taps.c1 = h_stage1[0];
    for (int i = 0; i < k1; ++i) taps.g1[i] = h_stage1[1 + i];
    taps.c2 = h_stage2[0];
    for (int i = 0; i < k2; ++i) taps.g2[i] = h_stage2[1 + i];
    const bool vec = aligned16(d_x) && aligned16(d_y) && ldx % 4 == 0 && ldy % 4 == 0;
    const size_t smem = (size_t)(kHbXS + kHbYS) * sizeof(float4);
    ECOG_TRY((smem_attr<halfband2_decimate_kernel>(smem)));
    dim3 grid((unsigned)ceil_div(ceil_div(T / 4, (int64_t)kHbM2), (int64_t)kHbTilesPerCta), (unsigned)C);
    halfband2_decimate_kernel<<<grid, kHbThreads, smem, (cudaStream_t)stream>>>(d_x, d_y, T, T / 4, ldx, ldy, vec, taps);
    return check_launch("halfband2_decimate");
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
