/*
 * libecog_sm100.so -- C ABI of the B200 (sm_100a) ECoG hot path.
 *
 * The reference (Daniel-Lin-S/decode_tonal_langauge) has no native layer: its hot
 * path is numpy/scipy called from Python plug-ins.  The functions below are what a
 * ctypes binding behind those plug-ins calls instead; each cites the reference
 * interface (file:line, relative to the reference root) whose arithmetic it replaces.
 * INTEGRATION.md shows the reference-side stub for every entry point.
 *
 * Conventions:
 *   - every pointer named d_* is DEVICE memory owned by the caller (PyTorch);
 *     h_* is HOST memory read during the call; the library never allocates,
 *     frees or synchronises the device and never creates streams;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*);
 *   - signals are float32, row-major (C channels, T samples) with a row stride
 *     `ld` in elements; statistics and filter states are float64;
 *   - return 0 on success or a negative ECOG_E_* code; ecog_last_error() gives
 *     the thread-local message.  The Python shims map ECOG_E_VALUE to ValueError
 *     (the reference's own error type for bad parameters) and the rest to
 *     RuntimeError;
 *   - no global mutable state: calls are re-entrant across host threads/streams.
 */
#ifndef ECOG_SM100_H
#define ECOG_SM100_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ECOG_ABI_VERSION 7

#define ECOG_OK 0
#define ECOG_E_VALUE (-1)     /* bad argument (reference raises ValueError)          */
#define ECOG_E_CUDA (-2)      /* CUDA launch / runtime failure                       */
#define ECOG_E_WORKSPACE (-3) /* workspace too small                                 */
#define ECOG_E_UNSUPPORTED (-4) /* shape outside what this build implements          */

typedef void* ecog_stream_t;

int ecog_abi_version(void);
const char* ecog_last_error(void);
/* number of kernels this library has launched from the calling thread (bench bookkeeping) */
int64_t ecog_launch_count(void);
/* comma-separated names of the kernels launched from the calling thread since the last reset (up to 4 KB);
 * reset != 0 clears the log after copying it.  Test / bench bookkeeping: which code path actually ran. */
const char* ecog_launch_log(int reset);

/* ------------------------------------------------------------------ K1: CAR
 * replaces preprocess/signal/car_rereference.py:34-39
 *   y[c,t] = x[c,t] - (1/n_inc) * sum_c w[c] x[c,t];  d_w (C floats, 0/1) may be NULL.
 * ecog_car is the single-GPU fused pass.  ecog_car_colsum / ecog_car_apply are the
 * two-phase form for channel-sharded recordings: the caller all-reduces d_colsum
 * (T floats) between the two enqueues.  Input and output rows have their own strides
 * (ldx, ldy >= T), like every other kernel family.                                  */
int ecog_car(const float* d_x, float* d_y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
             const float* d_w, double inv_count, ecog_stream_t stream);
int ecog_car_colsum(const float* d_x, int64_t C, int64_t T, int64_t ld, const float* d_w,
                    float* d_colsum, ecog_stream_t stream);
int ecog_car_apply(const float* d_x, float* d_y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                   const float* d_colsum, double inv_count, ecog_stream_t stream);

/* -------------------------------------------------------- K2: row statistics
 * replaces preprocess/signal/channel_zscore.py:22-27 and
 *          preprocess/signal/zscore_rereference.py:66-68
 * mean and POPULATION std (ddof=0) of x[c, t0:t1] in float64, then
 * y = (x - mean) / std over the whole row; nan_to_zero mirrors preserve_nans=False. */
size_t ecog_row_stats_workspace(int64_t C, int64_t T);
int ecog_row_stats(const float* d_x, int64_t C, int64_t T, int64_t ld, int64_t t0, int64_t t1,
                   double* d_mean, double* d_std, void* d_workspace, size_t workspace_bytes,
                   ecog_stream_t stream);
int ecog_zscore_apply(const float* d_x, float* d_y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                      const double* d_mean, const double* d_std, int nan_to_zero,
                      ecog_stream_t stream);

/* ---------------------------------------------- K3: IIR biquad cascade (SOS)
 * replaces preprocess/signal/frequency_filter.py:218-229 (scipy filtfilt / sosfilt).
 * h_sos: nsec x 6 float64 (b0 b1 b2 a0 a1 a2), host.  zero_phase=1 reproduces
 * filtfilt(padtype='odd', padlen): h_zi (nsec x 2, host) is the unit-step state in
 * cascade coordinates, scaled by the first padded sample of each sweep; zero_phase=0
 * is the causal zero-state sosfilt.  Rows are cut into chunks of `chunk` samples that
 * are filtered in parallel; h_M (2nsec x 2nsec, row major, host) is the cascade's
 * state-transition matrix raised to the power `chunk`, used to carry the state across
 * chunks (may be NULL when T <= chunk).  `tail` = number of trailing samples of a
 * chunk whose zero-state response still reaches the chunk end in float64
 * (ceil(log(1e-18)/log(max pole radius)), clamped to chunk).  d_y may alias d_x.     */
typedef struct {
    int32_t nsec;        /* biquad sections, 1..ECOG_MAX_SECTIONS                    */
    int32_t zero_phase;  /* 1 = forward-backward with odd padding, 0 = causal        */
    int32_t padlen;      /* filtfilt pad (27 for order-4 band filters); 0 if causal   */
    int32_t chunk;       /* samples per scan chunk, multiple of 16                    */
    int32_t tail;        /* multiple of 16; <= chunk in scan mode                     */
    int32_t mode;        /* ECOG_SOS_SCAN (exact carry scan) or ECOG_SOS_WARMUP        */
    int32_t threads;     /* warm-up mode: threads per CTA, 256 or 512; with ECOG_SOS_SPLIT_F32B: chunks per CTA,
                            a multiple of 32 in 32..256 (the CTA runs twice as many threads)            */
    int32_t split;       /* 0, or 4 (| ECOG_SOS_SPLIT_F32B): h_sos holds TWO 4-section unit-form cascades */
    int32_t tail_b;      /* split: warm-up samples of the second cascade (<= tail)     */
} ecog_sos_plan;
/* Cascade pair (split = 4, warm-up mode, zero phase, nsec = 8): two consecutive filtfilt steps
 * of the reference chain (e.g. the 58-62 Hz band-stop and the 70-150 Hz band-pass,
 * frequency_filter.py:218-229 twice) run as ONE forward and ONE backward sweep.  Away from the
 * row ends forward and backward sweeps of different LTI filters commute, so the result equals
 * the two sequential filtfilts there (to the warm-up bound); within `tail` samples of a row end
 * it does not, and the caller overwrites those samples with the sequential result computed on
 * short edge segments (decode_tonal_langauge_b200/ops.py::sosfilt_pair).  Sections 0-3 and 4-7
 * must each be of unit form (b2 = +-b0, all b0 = 1 except section 0, which carries the product
 * of both gains).                                                                           */
/* ECOG_SOS_WARMUP: one kernel per sweep; every chunk re-creates its start state by filtering
 * the `tail` samples before it from a zero state (valid when the host has checked that the
 * cascade's zero-input response is < 1e-10 after `tail` samples); chunks within `tail` of the
 * row edge use the exact filtfilt start-up instead.  h_M is not used.  The workspace then
 * also holds the forward result (C x T float32) of a zero-phase call.                       */
#define ECOG_SOS_SCAN 0
#define ECOG_SOS_WARMUP 1
/* ECOG_SOS_WARMUP_TMA: the warm-up sweeps with TMA-staged time tiles (cp.async.bulk.tensor boxes of
 * 256 chunks x 32 samples, SWIZZLE_128B, mbarrier completion; csrc/sosfilt_tma.cu).  Needs zero phase,
 * contiguous rows (ldx == ldy == T), T == n * chunk, chunk / tail (/ tail_b) multiples of 32,
 * tail <= chunk; 4-section cascades and the (2,5) / (0,5) cascade pairs.                        */
#define ECOG_SOS_WARMUP_TMA 2
/* ECOG_SOS_SPLIT_F32B (or-ed into `split`, TMA mode only): the SECOND cascade of a pair -- four (1 - z^-2)
 * band-pass sections -- runs in float32 delta form (states w[n-1] and w[n-1] - w[n-2], coefficients
 * -(1 + a1 + a2) and 1 - a2 rounded to float32 from the float64 design) while the first (the notch, pole
 * radius ~0.998) stays in float64: 12 of the pair's 29 FP64 operations per sample leave the FP64 pipe that
 * bounds the sweeps.  The caller asks for it only when the band-pass poles are far enough inside the unit
 * circle (min(1 - a2) >= 0.03: error <= ~1e-6 of the row maximum, measured 3.5e-7 / 4.5e-7 for 70-150 Hz at
 * 2 / 3 kHz against the all-float64 pair; decode_tonal_langauge_b200/design.py::bandpass_f32_ok).       */
#define ECOG_SOS_SPLIT_F32B 0x100
#define ECOG_MAX_SECTIONS 8
size_t ecog_sos_workspace(const ecog_sos_plan* plan, int64_t C, int64_t T);
int ecog_sosfilt(const float* d_x, float* d_y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                 const ecog_sos_plan* plan, const double* h_sos, const double* h_zi,
                 const double* h_M, void* d_workspace, size_t workspace_bytes,
                 ecog_stream_t stream);

/* strided device-to-device copy of a float32 (rows, cols) view (cudaMemcpy2DAsync on `stream`):
 * gathers / scatters the row-edge segments of the cascade pair without a host round trip.   */
int ecog_copy2d(const float* d_src, int64_t ld_src, float* d_dst, int64_t ld_dst, int64_t rows,
                int64_t cols, ecog_stream_t stream);

/* ------------------------------------ K4: Gaussian-Hilbert envelope (block-wise)
 * replaces preprocess/signal/frequency_filter.py:154-184.
 * Overlap-save with 4096-point shared-memory FFTs and a circular halo of `halo` samples.
 * Band b uses bins [h_shift[b], h_shift[b] + rows*256) of the block's half spectrum;
 * d_gain is nbands x (rows*256) float32 = Gaussian x analytic mask / (4096 * nbands) on
 * those bins (host, float64 -> float32).  rows in {1,2,4,8}.  With envelope=0 (real part)
 * every h_shift must be 0.  h_nz (optional, NULL = 16 per band; used when rows == 1): the
 * caller's promise that d_gain[b][k] == 0 for k >= 16 * h_nz[b] -- the inverse transforms then
 * skip the zero bins.  d_twiddle: table from ecog_hilbert_twiddles.
 * d_colsum (optional, NULL = none): T column sums from ecog_car_colsum; the kernel then filters
 * x[c,t] - inv_count * d_colsum[t], i.e. car_rereference.py:34-39 folded into the load (the
 * common-average step directly in front of the bank costs one read-only pass instead of a
 * read + write of the recording).                                                        */
#define ECOG_HILBERT_N 4096
size_t ecog_hilbert_twiddle_floats(void);
int ecog_hilbert_twiddles(float* h_out);   /* host helper: fills the per-thread twiddle table */
int ecog_hilbert_env(const float* d_x, float* d_y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                     const float* d_gain, int32_t nbands, int32_t rows, const int32_t* h_shift,
                     const int32_t* h_nz, int32_t halo, int32_t envelope, const float* d_twiddle,
                     const float* d_colsum, double inv_count, ecog_stream_t stream);
/* The same for overlap-save blocks [block_begin, block_end) only (block_begin even; block_end < 0 = all;
 * ecog_hilbert_blocks(T, halo) = number of blocks of a row; block b produces samples
 * [b U, (b+1) U) from inputs [b U - halo, (b+1) U + halo) circularly, U = 4096 - 2 halo).  Lets a
 * channel-sharded caller start the bank on the time range whose CAR column sums have already been
 * all-reduced while the collective still runs on the rest (distributed.py).                    */
int64_t ecog_hilbert_blocks(int64_t T, int32_t halo);
int ecog_hilbert_env_range(const float* d_x, float* d_y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                           const float* d_gain, int32_t nbands, int32_t rows, const int32_t* h_shift,
                           const int32_t* h_nz, int32_t halo, int32_t envelope, const float* d_twiddle,
                           const float* d_colsum, double inv_count, int64_t block_begin, int64_t block_end,
                           ecog_stream_t stream);

/* ------------------------------------------------- K5: whole-row FFT resample
 * replaces preprocess/signal/downsample.py:21-27 (scipy.signal.resample, real input).
 * Plan arrays are built by the host (decode_tonal_langauge_b200/fftplan.py).          */
typedef struct {
    int32_t n;           /* FFT length along the tile axis                           */
    int32_t nstage;      /* radix stages                                             */
    int32_t radix[16];   /* product == n, each in {2,3,4,5}                          */
} ecog_fft_axis;
typedef struct {
    int64_t T;           /* input samples per row (even)                             */
    int64_t num;         /* output samples per row (even)                            */
    ecog_fft_axis fa, fb;   /* forward: N = T/2 = fa.n * fb.n (column pass, row pass)  */
    ecog_fft_axis ia, ib;   /* inverse: N' = num/2 = ia.n * ib.n                       */
} ecog_resample_plan;
/* device tables, all built on the host in float64 and rounded once:
 *   d_tw_*    float2 roots of unity W_n^k (length = axis n)
 *   d_tw_big_f / d_tw_big_i  float2 four-step twiddle factors, two-level tables
 *   d_untangle float2 x (num/2 + 1) pairs used by the spectrum repack                 */
typedef struct {
    const float *tw_fa, *tw_fb, *tw_ia, *tw_ib;
    const float *tw_big_f_hi, *tw_big_f_lo, *tw_big_i_hi, *tw_big_i_lo;
    int32_t big_f_split, big_i_split;
    const float *tw_T, *tw_num;
    const float *bin_gain;   /* optional (NULL = none): num/2 + 1 per-bin gains applied to the kept
                                spectrum, i.e. 1 / H1[k] of the FIR pre-decimator (ecog_fir_decimate) */
    const double *tw_q_f, *tw_q_i;   /* double2 W_N^q, q < fa.n (resp. W_N'^q, q < ia.n): float64 base of
                                        the column powers of the four-step twiddle                      */
} ecog_resample_tables;
size_t ecog_resample_workspace(const ecog_resample_plan* plan, int64_t C);
int ecog_fft_resample(const float* d_x, float* d_y, int64_t C, int64_t ldx, int64_t ldy,
                      const ecog_resample_plan* plan, const ecog_resample_tables* tables,
                      void* d_workspace, size_t workspace_bytes, ecog_stream_t stream);

/* ----------------------------- K5b: building blocks of the chirp-z (Bluestein) path
 * downsample.py:21-27 for rows whose length is NOT 2-3-5 smooth (real TDT rates: T = 1 831 054,
 * num = 239 999): X[k] = conj(w[k]) (x conj(w) * w)[k] with w[n] = exp(i pi n^2 / T), the
 * convolution done with smooth-length complex FFTs; same for the inverse of `num` points.
 * Host glue: decode_tonal_langauge_b200/fftplan.py::czt_plan, ops._czt_resample.
 *   ecog_fft_c2c      : C rows of N = fa.n * fb.n complex points (interleaved float), natural
 *                       order in and out; inverse = conj -> forward -> conj; out *= scale.
 *                       Strides in complex elements; workspace from ecog_fft_c2c_workspace.
 *   ecog_cplx_modulate: out[c,i] = (i < in_len ? in[c,i] * table[i] : 0), i < out_len (table has
 *                       in_len complex entries), with real
 *                       or complex input and complex or real-part output (zero padding, chirp
 *                       and spectrum multiplications).  May run in place.                      */
typedef struct {
    const float *tw_a, *tw_b;            /* W_n^k per axis                                    */
    const float *tw_big_hi, *tw_big_lo;  /* four-step twiddles, two-level table (split 4096)  */
    const double *tw_q;                  /* double2 W_N^q, q < fa.n                           */
} ecog_fft_tables;
size_t ecog_fft_c2c_workspace(const ecog_fft_axis* fa, const ecog_fft_axis* fb, int64_t C);
int ecog_fft_c2c(const float* d_in, float* d_out, int64_t C, int64_t ld_in, int64_t ld_out,
                 const ecog_fft_axis* fa, const ecog_fft_axis* fb, const ecog_fft_tables* tables,
                 int32_t inverse, float scale, void* d_workspace, size_t workspace_bytes,
                 ecog_stream_t stream);
int ecog_cplx_modulate(const float* d_in, int32_t in_is_complex, int64_t in_len, int64_t ld_in,
                       const float* d_table /* NULL = 1 */, float* d_out, int32_t out_is_complex,
                       int64_t out_len, int64_t ld_out, int64_t C, ecog_stream_t stream);
/* acc[c,i] (+)= scale * |z[c,i]| (envelope) or scale * Re z[c,i]: band accumulation of the
 * whole-record Gaussian-Hilbert path (frequency_filter.py:171-184) used for banks whose time
 * kernels do not fit the 4096-sample block of ecog_hilbert_env (low-frequency bands).           */
int ecog_cplx_abs_accumulate(const float* d_z, int64_t ld_z, float* d_acc, int64_t ld_acc, int64_t len,
                             int64_t C, int32_t envelope, float scale, int32_t accumulate,
                             ecog_stream_t stream);

/* ------------------------------- K5a: circular FIR low-pass + integer decimation
 * First stage of the two-stage realisation of downsample.py:21-27 for large ratios:
 *   y[c, m] = sum_{j<ntaps} h[j] * x[c, (m*D + j - offset) mod T],  m = 0 .. T/D - 1.
 * The brick wall of scipy.signal.resample is then applied by ecog_fft_resample on the T/D
 * samples per row, with bin_gain[k] = 1 / H1[k] dividing this stage's (known, non-zero)
 * pass-band response out exactly; bins k <= num/2 are alias-free to the stop-band
 * attenuation of h (120 dB Kaiser design, decode_tonal_langauge_b200/fftplan.py).
 * h_taps: host, ntaps <= 256.  D in {2, 4}; T % D == 0.                               */
int ecog_fir_decimate(const float* d_x, float* d_y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                      const float* h_taps, int32_t ntaps, int32_t offset, int32_t D,
                      ecog_stream_t stream);
/* Decimation by 4 as TWO circular half-band stages (the same role in downsample.py:21-27, ABI 7): a decimate-by-2
 * stage in front of a brick wall at f_pass has the band edges f_pass and 1 - f_pass, symmetric about half the
 * Nyquist frequency, so every second tap of both stages is zero:
 *   y1[c, n] = h1[0] x[c, 2n] + sum_{i<k1} h1[1+i] (x[c, (2n-2i-1) mod T] + x[c, (2n+2i+1) mod T]),   n < T/2
 *   y [c, m] = h2[0] y1[c, 2m] + sum_{i<k2} h2[1+i] (y1[c, (2m-2i-1) mod T/2] + y1[c, (2m+2i+1) mod T/2]),  m < T/4
 * (19.25 instead of 40 products per input sample for 2 kHz -> 400 Hz; y1 never leaves shared memory).  The FFT stage
 * then divides bin k by H1[k] H2[k] = (h1[0] + 2 sum h1[1+i] cos(2 pi k (2i+1) / T)) (h2[0] + 2 sum h2[1+i] cos(4 pi k (2i+1) / T)).
 * h_stage1: 1 + k1 floats (k1 <= 8), h_stage2: 1 + k2 floats (k2 <= 21), host.  T % 4 == 0.           */
int ecog_halfband2_decimate(const float* d_x, float* d_y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                            const float* h_stage1, int32_t k1, const float* h_stage2, int32_t k2,
                            ecog_stream_t stream);

/* ------------------------------------------------------- K6: causal FIR (long)
 * replaces preprocess/signal/frequency_filter.py:260-274 (`fir` band method).  The mean over
 * centre frequencies of lfilter(firwin(order+1, ...), 1, x) is one causal FIR with the
 * averaged taps h (zero initial state):  y[c,t] = sum_j h[j] x[c,t-j].
 * d_taps_rev: DEVICE, 4*ntaps4 floats, g[i] = h[off - i] (0 where off - i is outside h) with
 * off = 4*ntaps4 - 4 >= len(h) - 1, 16-byte aligned.                                    */
int ecog_fir_causal(const float* d_x, float* d_y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                    const float* d_taps_rev, int32_t ntaps4, ecog_stream_t stream);

/* ------------------------------------------------ K7: trailing-window z-score
 * replaces preprocess/signal/rolling_zscore.py:28-49 (pandas rolling(window, min_periods=1)
 * mean and std with ddof=1): z[t] = (x[t] - mean_w) / std_w over x[max(0,t+1-window) .. t];
 * sample 0 of every row is NaN (0 when nan_to_zero).  d_shift: per-row float64 offset
 * subtracted before the float64 prefix sums (the row mean from ecog_row_stats).          */
size_t ecog_rolling_workspace(int64_t C, int64_t T);
int ecog_rolling_zscore(const float* d_x, float* d_y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                        int64_t window, const double* d_shift, int nan_to_zero,
                        void* d_workspace, size_t workspace_bytes, ecog_stream_t stream);

/* ------------------------------------------------------- K8: epoch gather
 * replaces data_loading/text_align.py:290-304,331-340,380-394.
 *   out[n, c, 0:L] = src[c, start[n] : start[n]+L]    (bit copy; elem_bytes 1, 2, 4 or 8)
 * d_start: int64 onset indices computed on the host in float64 (Appendix A6).
 * Returns ECOG_E_VALUE if any window leaves [0, T) -- checked on the host copy h_start. */
int ecog_epoch_gather(const void* d_src, void* d_out, int64_t C, int64_t T, int64_t ld,
                      const int64_t* d_start, const int64_t* h_start, int64_t N, int64_t L,
                      int32_t elem_bytes, ecog_stream_t stream);

/* ------------------------------------------- K8b: channel selection of epochs
 * replaces data_loading/sample_loading.py:77-79 (features[:, channels, :], the consumer of the
 * channel-selection JSON): out[n, j, 0:L] = src[n, channels[j], 0:L], bit copy.            */
int ecog_channel_select(const void* d_src, void* d_out, int64_t N, int64_t C, int64_t L,
                        const int32_t* d_channels, const int32_t* h_channels, int64_t K,
                        int32_t elem_bytes, ecog_stream_t stream);

/* --------------------------------------------- K9/K10: ANOVA F + run length
 * replaces channel_selection/discriminative.py:172-180, channel_selection/active.py:58-76,
 *          channel_selection/utils.py:4-30,63-75 (scipy.stats.f_oneway, equal_var).
 * Epochs may come from up to two tensors (ERP then rest, for active.run, numbered 0..Na-1
 * then Na..Na+Nb-1); d_order lists the event numbers SORTED BY GROUP (stable argsort of the
 * group index 0..G-1) and h_group_count[k] says how many belong to group k.
 * Outputs F and p = fdtrc(G-1, N-G, F) as (C, L) float64.                                */
size_t ecog_anova_workspace(int64_t C, int64_t L, int64_t N, int32_t G);
int ecog_anova_f(const float* d_epochs_a, int64_t Na, const float* d_epochs_b, int64_t Nb,
                 int64_t C, int64_t L, const int32_t* d_order, const int64_t* h_group_count,
                 int32_t G, double* d_F, double* d_p, void* d_workspace, size_t workspace_bytes,
                 ecog_stream_t stream);
/* longest run of consecutive p < threshold per channel (NaN compares false) */
int ecog_sig_runlength(const double* d_p, int64_t C, int64_t L, double threshold,
                       int32_t* d_maxrun, ecog_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* ECOG_SM100_H */
