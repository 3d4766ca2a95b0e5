// K4 Gaussian filter-bank analytic envelope ("hilbert" method), block-wise.
//
// The reference takes one FFT of the whole record, multiplies by nb Gaussian x
// analytic-mask kernels and inverse-transforms each (frequency_filter.py:154-184).
// Every band kernel is a Gaussian in time with sigma_t <= a few tens of ms, so the
// circular convolution is reproduced exactly (to < 1e-9) by overlap-save on 4096-sample
// blocks with a `halo` of >= 6.5 sigma_t on each side, wrapped circularly at the
// record ends.  One CTA (256 threads) processes TWO consecutive blocks of one channel:
//   - the two real blocks ride as real/imag parts of ONE forward complex FFT and are
//     separated with the conjugate-symmetry split;
//   - per band and block: multiply by the gain table, inverse FFT (as conj-forward),
//     |.| (or real part) accumulated in registers -> mean over bands never leaves the SM;
//   - a band occupies only a few hundred of the 2048 positive bins.  For the envelope the
//     band is shifted down to bin 0 (|z| is invariant to a spectral shift), so only
//     `rows` x 256 inputs of the inverse transform are non-zero and its first radix-16
//     pass collapses to `rows` terms (rows = 1 for the high-gamma bank at 2 kHz).
// FFT: 4096 = 16 x 16 x 16, each thread holds 16 points in registers, three radix-16
// passes with two shared-memory exchanges (padded index i + i/16, conflict free).
// The kernel is FP32-ALU/shared-memory bound (~9 FFTs per 4096 samples), NOT HBM bound:
// algorithmic traffic is 8 B per sample (read x once + halo, write y once).
#include "common.cuh"

namespace ecog {

constexpr int kN = ECOG_HILBERT_N;      // 4096
constexpr int kHT = 256;                // threads
constexpr int kHalf = kN / 2;

__device__ __forceinline__ int padi(int i) { return i + (i >> 4); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(fmaf(a.x, b.x, -a.y * b.y), fmaf(a.x, b.y, a.y * b.x));
}
#ifndef ECOG_SCALAR_F32
// complex add / sub as ONE packed instruction on the (re, im) register pair (sm_100 add.f32x2 / sub.f32x2)
__device__ __forceinline__ float2 cadd(float2 a, float2 b) {
    float2 r;
    asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; add.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
__device__ __forceinline__ float2 csub(float2 a, float2 b) {
    float2 r;
    asm("{ .reg .b64 ra, rb, rc; mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; sub.rn.f32x2 rc, ra, rb; mov.b64 {%0, %1}, rc; }"
        : "=f"(r.x), "=f"(r.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
    return r;
}
#else
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
#endif
// multiply by -i (forward-DFT quarter turn)
__device__ __forceinline__ float2 mul_mi(float2 a) { return make_float2(a.y, -a.x); }

// forward 4-point DFT in place: (a0,a1,a2,a3) -> X[0..3]
__device__ __forceinline__ void dft4(float2& a0, float2& a1, float2& a2, float2& a3) {
    float2 t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = mul_mi(csub(a1, a3));
    a0 = cadd(t0, t2); a2 = csub(t0, t2); a1 = cadd(t1, t3); a3 = csub(t1, t3);
}

// 4-point DFT of (x0, w1 y1, w2 y2, w3 y3) with the twiddles w_m = s_m (1 + i tau_m) folded into
// the butterflies: z_m = y_m (1 + i tau_m) costs two FMAs, and every scale s_m rides in an FMA that
// would have been an addition (rho = s3 / s1):  22 operations instead of 28 (3 complex products +
// 8 complex additions).  NEGI2: w2 = -i exactly (no scale).
template <bool NEGI2>
__device__ __forceinline__ void dft4_tw(float2& x0, float2& y1, float2& y2, float2& y3,
                                        float tau1, float s1, float tau2, float s2, float tau3, float rho) {
    const float2 z1 = make_float2(fmaf(-tau1, y1.y, y1.x), fmaf(tau1, y1.x, y1.y));
    const float2 z3 = make_float2(fmaf(-tau3, y3.y, y3.x), fmaf(tau3, y3.x, y3.y));
    float2 t0, t1;
    if (NEGI2) {
        const float2 x2 = make_float2(y2.y, -y2.x);
        t0 = cadd(x0, x2); t1 = csub(x0, x2);
    } else {
        const float2 z2 = make_float2(fmaf(-tau2, y2.y, y2.x), fmaf(tau2, y2.x, y2.y));
        t0 = make_float2(fmaf(s2, z2.x, x0.x), fmaf(s2, z2.y, x0.y));
        t1 = make_float2(fmaf(-s2, z2.x, x0.x), fmaf(-s2, z2.y, x0.y));
    }
    const float2 u = make_float2(fmaf(rho, z3.x, z1.x), fmaf(rho, z3.y, z1.y));
    const float2 d = make_float2(fmaf(-rho, z3.x, z1.x), fmaf(-rho, z3.y, z1.y));
    x0 = make_float2(fmaf(s1, u.x, t0.x), fmaf(s1, u.y, t0.y));
    y2 = make_float2(fmaf(-s1, u.x, t0.x), fmaf(-s1, u.y, t0.y));
    y1 = make_float2(fmaf(s1, d.y, t1.x), fmaf(-s1, d.x, t1.y));          // t1 + s1 (-i d)
    y3 = make_float2(fmaf(-s1, d.y, t1.x), fmaf(s1, d.x, t1.y));
}

// second half of the 16-point DFT: inter-stage twiddles W16^(m p) + the four DFT4 over m + the
// transposition to natural order; element (m, p) lives at v[m + 4p] on entry
__device__ __forceinline__ void dft16_stage2(float2 (&v)[16]) {
    const float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, C2 = 0.70710678118654752f;
    const float T1 = 0.41421356237309503f, T3 = 2.4142135623730951f;    // tan(pi/8), tan(3 pi/8)
    dft4(v[0], v[1], v[2], v[3]);
    // p = 1: W^1 = C1 (1 - i T1), W^2 = C2 (1 - i), W^3 = S1 (1 - i T3)
    dft4_tw<false>(v[4], v[5], v[6], v[7], -T1, C1, -1.f, C2, -T3, S1 / C1);
    // p = 2: W^2 = C2 (1 - i), W^4 = -i, W^6 = -C2 (1 + i)
    dft4_tw<true>(v[8], v[9], v[10], v[11], -1.f, C2, 0.f, 0.f, 1.f, -1.f);
    // p = 3: W^3 = S1 (1 - i T3), W^6 = -C2 (1 + i), W^9 = -C1 (1 - i T1)
    dft4_tw<false>(v[12], v[13], v[14], v[15], -T3, S1, 1.f, -C2, -T1, -C1 / S1);
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int r = p + 1; r < 4; ++r) { float2 t = v[4 * p + r]; v[4 * p + r] = v[4 * r + p]; v[4 * r + p] = t; }
}

// forward 16-point DFT in place, natural order in and out
__device__ __forceinline__ void dft16(float2 (&v)[16]) {
    // stage 1: for each m, DFT4 over q of v[m + 4q]  -> a[m][p] kept at v[m + 4p]
#pragma unroll
    for (int m = 0; m < 4; ++m) dft4(v[m], v[m + 4], v[m + 8], v[m + 12]);
    // twiddles W16^(m p) + stage 2 (for each p, DFT4 over m of v[m + 4p] -> X[p + 4r]) + transposition
    dft16_stage2(v);
}

// passes 2 and 3 of the 4096-point forward FFT; pass-1 output must already be in `buf`
// (index padi(256 k0 + tid)).  Result: v[k2] = X[k0 + 16 k1 + 256 k2] with tid = 16 k0 + k1.
__device__ __forceinline__ void fft4096_finish(float2 (&v)[16], float2* buf, const float2* __restrict__ tw2, int tid) {
    float2 t2[16];
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) t2[k1] = __ldg(&tw2[k1 * kHT + tid]);   // in flight across the barrier
    __syncthreads();
    const int k0 = tid >> 4, n0 = tid & 15;
    float2* p2 = buf + 272 * k0 + n0;            // padi(256 k0 + 16 n1 + n0) = 272 k0 + 17 n1 + n0
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) v[n1] = p2[17 * n1];
    dft16(v);
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) v[k1] = cmul(v[k1], t2[k1]);
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) p2[17 * k1] = v[k1];
    __syncwarp();                                // this exchange stays inside the half-warp that shares k0
    const float2* p3 = buf + 17 * tid;           // padi(16 tid + j) = 17 tid + j
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = p3[j];
    dft16(v);
}

// pass 1: v[n2] = x[256 n2 + tid] -> twiddled A[k0] stored at padi(256 k0 + tid)
__device__ __forceinline__ void fft4096_pass1(float2 (&v)[16], float2* buf, const float2* __restrict__ tw1, int tid) {
    dft16(v);
#pragma unroll
    for (int k0 = 1; k0 < 16; ++k0) v[k0] = cmul(v[k0], __ldg(&tw1[k0 * kHT + tid]));
    float2* p1 = buf + tid + (tid >> 4);         // padi(256 k0 + tid) = 272 k0 + tid + tid/16
#pragma unroll
    for (int k0 = 0; k0 < 16; ++k0) p1[272 * k0] = v[k0];
}

// First radix-16 pass of an inverse transform whose input is non-zero only in rows
// n2 < CNT (k = 256 n2 + tid): the band occupies CNT*256 consecutive bins after the
// per-band spectral shift.  CNT = 1: A[k0] = v0;  2: v0 + W16^k0 v1;  4 / 8: the first
// radix-4 layer of dft16 degenerates to copies / 2-point butterflies.
template <int CNT>
__device__ __forceinline__ void dft16_rows(float2 (&v)[16]) {
    if (CNT == 1) {
#pragma unroll
        for (int k = 1; k < 16; ++k) v[k] = v[0];
        return;
    }
    if (CNT == 2) {
        const float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, C2 = 0.70710678118654752f;
        const float2 a = v[0], b = v[1];
        const float wr[16] = {1.f, C1, C2, S1, 0.f, -S1, -C2, -C1, -1.f, -C1, -C2, -S1, 0.f, S1, C2, C1};
        const float wi[16] = {0.f, -S1, -C2, -C1, -1.f, -C1, -C2, -S1, 0.f, S1, C2, C1, 1.f, C1, C2, S1};
#pragma unroll
        for (int k = 0; k < 16; ++k)
            v[k] = make_float2(fmaf(wr[k], b.x, fmaf(-wi[k], b.y, a.x)), fmaf(wr[k], b.y, fmaf(wi[k], b.x, a.y)));
        return;
    }
    const float C1 = 0.92387953251128674f, S1 = 0.38268343236508977f, C2 = 0.70710678118654752f;
    // stage 1: a[m][p] = sum_q v[m + 4q] W4^{qp}, only q < CNT/4 present; a[m][p] kept at v[m + 4p]
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        if (CNT == 4) {
            v[m + 4] = v[m]; v[m + 8] = v[m]; v[m + 12] = v[m];
        } else {   // CNT == 8: q in {0, 1}
            const float2 x0 = v[m], x1 = v[m + 4];
            v[m] = cadd(x0, x1); v[m + 8] = csub(x0, x1);
            v[m + 4] = cadd(x0, mul_mi(x1)); v[m + 12] = csub(x0, mul_mi(x1));
        }
    }
    v[1 + 4] = cmul(v[1 + 4], make_float2(C1, -S1));
    v[1 + 8] = cmul(v[1 + 8], make_float2(C2, -C2));
    v[1 + 12] = cmul(v[1 + 12], make_float2(S1, -C1));
    v[2 + 4] = cmul(v[2 + 4], make_float2(C2, -C2));
    v[2 + 8] = mul_mi(v[2 + 8]);
    v[2 + 12] = cmul(v[2 + 12], make_float2(-C2, -C2));
    v[3 + 4] = cmul(v[3 + 4], make_float2(S1, -C1));
    v[3 + 8] = cmul(v[3 + 8], make_float2(-C2, -C2));
    v[3 + 12] = cmul(v[3 + 12], make_float2(-C1, S1));
#pragma unroll
    for (int p = 0; p < 4; ++p) dft4(v[4 * p], v[4 * p + 1], v[4 * p + 2], v[4 * p + 3]);
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int r = p + 1; r < 4; ++r) { float2 t = v[4 * p + r]; v[4 * p + r] = v[4 * r + p]; v[4 * r + p] = t; }
}

__device__ __forceinline__ float fast_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

struct BandShift { int s[64]; int nz[64]; };   // spectral shift and non-zero 16-bin groups per band

template <int CNT>
__global__ void __launch_bounds__(kHT, 2)
hilbert_env_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t T, int64_t ldx, int64_t ldy,
                   const float* __restrict__ gain, int nb, BandShift shift, int halo, int envelope,
                   const float2* __restrict__ tw, int64_t nBlocks, const float* __restrict__ colsum, float inv_count,
                   int64_t blk_begin, int64_t blk_end) {
    extern __shared__ __align__(16) float2 hsm[];
    float2* bufA = hsm;                       // [4096 + 256]  conj spectra of block 0 | block 1 (2048 each)
    float2* bufB = hsm + (kN + kN / 16);      // [4096 + 256]  exchange buffer
    const int tid = threadIdx.x;
    const int64_t ch = blockIdx.y;
    const int U = kN - 2 * halo;
    const int64_t b0 = blk_begin + 2 * (int64_t)blockIdx.x, b1 = b0 + 1;
    const float* xr = x + ch * ldx;
    const float2* tw1 = tw;                 // [16][256] : W_256^{(tid>>4) k0}
    const float2* tw2 = tw + 16 * kHT;      // [16][256] : W_4096^{(tid&15)((tid>>4) + 16 k1)}
    const int k0 = tid >> 4, k1 = tid & 15;
    const int nat0 = k0 + 17 * k1;          // padi(k0 + 16 k1 + 256 k2) = nat0 + 272 k2
    const int p1 = tid + k0;                // padi(256 r + tid)         = p1 + 272 r

    float2 v[16];
    {   // two real blocks as one complex signal, circular halo
        int64_t s0 = (b0 * U - halo) % T; if (s0 < 0) s0 += T;
        int64_t s1 = (b1 * U - halo) % T; if (s1 < 0) s1 += T;
        const bool has1 = b1 < blk_end;
        const bool wrap = (s0 + kN > T) || (s1 + kN > T);
        if (!wrap) {
#pragma unroll
            for (int n2 = 0; n2 < 16; ++n2) {
                const int i = 256 * n2 + tid;
                v[n2].x = xr[s0 + i];
                v[n2].y = has1 ? xr[s1 + i] : 0.f;
            }
            if (colsum) {     // common-average reference folded into the load: x - (1/n) sum_c w_c x_c
#pragma unroll
                for (int n2 = 0; n2 < 16; ++n2) {
                    const int i = 256 * n2 + tid;
                    v[n2].x = fmaf(-inv_count, __ldg(colsum + s0 + i), v[n2].x);
                    if (has1) v[n2].y = fmaf(-inv_count, __ldg(colsum + s1 + i), v[n2].y);
                }
            }
        } else {
#pragma unroll
            for (int n2 = 0; n2 < 16; ++n2) {
                const int i = 256 * n2 + tid;
                const int64_t i0 = (s0 + i) % T, i1 = (s1 + i) % T;
                v[n2].x = xr[i0];
                v[n2].y = has1 ? xr[i1] : 0.f;
                if (colsum) {
                    v[n2].x = fmaf(-inv_count, __ldg(colsum + i0), v[n2].x);
                    if (has1) v[n2].y = fmaf(-inv_count, __ldg(colsum + i1), v[n2].y);
                }
            }
        }
    }
    fft4096_pass1(v, bufA, tw1, tid);
    fft4096_finish(v, bufA, tw2, tid);
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) bufB[nat0 + 272 * k2] = v[k2];      // natural-order spectrum Z[k]
    __syncthreads();
    // conjugate-symmetry split (factor 1/2 folded into the gain table):
    //   block0: S0[k] = Z[k] + conj(Z[N-k]),  block1: S1[k] = -i (Z[k] - conj(Z[N-k]))
    // stored CONJUGATED (the inverse transforms run as conj-forward FFTs)
#pragma unroll
    for (int j = 0; j < kHalf / kHT; ++j) {
        const int k = tid + kHT * j;
        float2 s0 = make_float2(0.f, 0.f), s1 = make_float2(0.f, 0.f);
        if (k > 0) {
            const float2 zk = bufB[padi(k)], zm = bufB[padi(kN - k)];
            const float2 a = make_float2(zk.x + zm.x, zk.y - zm.y);     // Z[k] + conj(Z[N-k])
            const float2 d = make_float2(zk.x - zm.x, zk.y + zm.y);     // Z[k] - conj(Z[N-k])
            s0 = make_float2(a.x, -a.y);                                 // conj(a)
            s1 = make_float2(d.y, d.x);                                  // conj(-i d)
        }
        bufA[padi(k)] = s0;
        bufA[padi(kHalf + k)] = s1;
    }
    __syncthreads();

    float* yr = y + ch * ldy;
    for (int sel = 0; sel < 2; ++sel) {
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0.f;
        for (int band = 0; band < nb; ++band) {
            const float* g = gain + (size_t)band * (CNT * 256);
            const int sh = shift.s[band];
            // band spectrum shifted down by sh bins: rows n2 < CNT only (|z| is shift invariant;
            // sh = 0 when the real part is requested)
#pragma unroll
            for (int n2 = 0; n2 < CNT; ++n2) {
                const int j = 256 * n2 + tid;
                const float gk = __ldg(&g[j]);
                const int k = sh + j;                        // < 2048 by construction of the table
                const float2 sv = bufA[padi(sel * kHalf + k)];
                v[n2] = make_float2(sv.x * gk, sv.y * gk);
            }
            float2 t1[16];
#pragma unroll
            for (int e = 1; e < 16; ++e) t1[e] = __ldg(&tw1[e * kHT + tid]);
            dft16_rows<CNT>(v);
#pragma unroll
            for (int e = 1; e < 16; ++e) v[e] = cmul(v[e], t1[e]);
            __syncthreads();                       // previous transform's pass-3 reads of bufB are done
#pragma unroll
            for (int e = 0; e < 16; ++e) bufB[p1 + 272 * e] = v[e];
            fft4096_finish(v, bufB, tw2, tid);
            // v[k2] = conj(z[t]) (times a unit phasor when shifted), t = k0 + 16 k1 + 256 k2
            if (envelope) {
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] += fast_sqrt(fmaf(v[j].x, v[j].x, v[j].y * v[j].y));
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] += v[j].x;
            }
        }
        __syncthreads();                           // all pass-3 reads of bufB done
        float* ob = reinterpret_cast<float*>(bufB);   // natural order floats, pitch as padi
#pragma unroll
        for (int k2 = 0; k2 < 16; ++k2) ob[nat0 + 272 * k2] = acc[k2];
        __syncthreads();
        const int64_t bb = sel == 0 ? b0 : b1;
        if (bb < blk_end) {
            for (int i = tid; i < U; i += kHT) {
                const int64_t t = bb * U + i;
                if (t < T) yr[t] = ob[padi(halo + i)];
            }
        }
        __syncthreads();                           // ob is reused as the exchange buffer
    }
}

// ---------------------------------------------------------------- fast path: <= 8 narrow bands
// Specialisation for banks whose bands fit one row of 256 bins (the high-gamma bank at 2 kHz).
// The generic kernel is shared-memory-bandwidth bound: per inverse transform every thread moves
// 65 float2 through shared memory and 31 twiddles through L1.  Here
//   * the gained, shifted band spectra SG[block][band][256] are built ONCE per CTA (one complex
//     product per thread and band) straight from the natural-order spectrum;
//   * an inverse has only 256 non-zero inputs, so its first radix-16 pass is the identity: the
//     thread (k0, n0) of pass 2 reads its 16 inputs SG[16 n1 + n0] directly (broadcast loads,
//     one wavefront each) and multiplies by W_256^{n1 k0} W_4096^{n0 k0} held in registers;
//   * the remaining inter-pass twiddle is W_256^{n0 k1}: a 16 x 16 table in shared memory;
//   * the exchange buffer alternates between two regions, so one barrier per transform.
// Per inverse and thread: 16 + 16 + 16 + 16 shared-memory accesses instead of 65 + 31.
constexpr int kFastBands = 8;

// t0 = a wa + b wb, t1 = a wa - b wb with the products folded into the butterfly:
// p = a wa (4 ops), t0 = p + b wb (4 FMA), t1 = 2 p - t0 (2 FMA): 10 operations instead of 12.
__device__ __forceinline__ void tw_bfly(float2 a, float2 wa, float2 b, float2 wb, float2& t0, float2& t1) {
    const float2 p = cmul(a, wa);
    t0 = make_float2(fmaf(b.x, wb.x, fmaf(-b.y, wb.y, p.x)), fmaf(b.x, wb.y, fmaf(b.y, wb.x, p.y)));
    t1 = make_float2(fmaf(2.f, p.x, -t0.x), fmaf(2.f, p.y, -t0.y));
}
// same with wa = 1
__device__ __forceinline__ void tw_bfly1(float2 a, float2 b, float2 wb, float2& t0, float2& t1) {
    t0 = make_float2(fmaf(b.x, wb.x, fmaf(-b.y, wb.y, a.x)), fmaf(b.x, wb.y, fmaf(b.y, wb.x, a.y)));
    t1 = make_float2(fmaf(2.f, a.x, -t0.x), fmaf(2.f, a.y, -t0.y));
}

// forward 16-point DFT of v[i] * w[i], natural order in and out: the input twiddles ride in the
// first butterfly layer.  W0ONE: w[0] == 1.  NZ: inputs v[NZ..15] are zero (8 <= NZ <= 16) and are
// neither read nor multiplied.
template <bool W0ONE, int NZ>
__device__ __forceinline__ void dft16_tw(float2 (&v)[16], const float2 (&w)[16]) {
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        float2 t0, t1, t2, t3;
        if (m + 8 < NZ) {
            if (W0ONE && m == 0) tw_bfly1(v[0], v[8], w[8], t0, t1);
            else tw_bfly(v[m], w[m], v[m + 8], w[m + 8], t0, t1);
        } else {
            t0 = t1 = (W0ONE && m == 0) ? v[0] : cmul(v[m], w[m]);
        }
        if (m + 12 < NZ) {
            tw_bfly(v[m + 4], w[m + 4], v[m + 12], w[m + 12], t2, t3);
        } else {
            t2 = t3 = cmul(v[m + 4], w[m + 4]);
        }
        t3 = mul_mi(t3);
        v[m] = cadd(t0, t2); v[m + 8] = csub(t0, t2); v[m + 4] = cadd(t1, t3); v[m + 12] = csub(t1, t3);
    }
    dft16_stage2(v);
}

// pass 2 of an inverse whose gained band spectrum has NZ * 16 non-zero bins: thread (k0, n0) reads
// SG[n0][n1 < NZ] (128-bit loads) and transforms over n1 with W_256^{n1 k0} W_4096^{n0 k0} folded in
template <int NZ>
__device__ __forceinline__ void inverse_pass2(float2 (&v)[16], const float2* __restrict__ sgrow, const float2 (&twA)[16]) {
    const float4* sg = reinterpret_cast<const float4*>(sgrow);
#pragma unroll
    for (int i = 0; i < (NZ + 1) / 2; ++i) {
        const float4 q = sg[i];
        v[2 * i] = make_float2(q.x, q.y);
        v[2 * i + 1] = make_float2(q.z, q.w);
    }
    dft16_tw<false, NZ>(v, twA);
}

constexpr int kP18 = 18;                                  // pitch (float2) of 16-entry rows read with LDS.128
constexpr int kXchg = 16 * 16 * kP18;                     // 4608: exchange buffer of the inverse transforms
constexpr int kSgBand = 16 * kP18;                        // 288: one gained band spectrum, layout [n0][n1]
constexpr size_t kFastSmem = ((size_t)kXchg + (kN + kN / 16) + 2 * kFastBands * kSgBand + kSgBand) * sizeof(float2);

// EDGE: the output samples t < 256 and t >= 3840 of a block (k2 = 0 and 15) are needed, i.e. the
// halo is shorter than 256 samples; otherwise they are never computed.
template <bool ENV, bool EDGE>
__global__ void __launch_bounds__(kHT, 2)
hilbert_env8_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t T, int64_t ldx, int64_t ldy,
                    const float* __restrict__ gain, int nb, BandShift shift, int halo,
                    const float2* __restrict__ tw, int64_t nBlocks, const float* __restrict__ colsum, float inv_count,
                    int64_t blk_begin, int64_t blk_end, int chan_fast) {
    extern __shared__ __align__(16) float2 hsm[];
    float2* bufA = hsm;                                   // [4608] forward exchange buffer / inverse exchange
    float2* bufB = bufA + kXchg;                          // [4096 + 256] natural-order spectrum / output staging
    float2* SG = bufB + (kN + kN / 16);                   // [2][kFastBands][16 n0][18] gained band spectra (conjugated)
    float2* twBs = SG + 2 * kFastBands * kSgBand;         // [16 k1][18] W_256^{n0 k1}
    const int tid = threadIdx.x;
    // chan_fast: consecutive CTAs are the SAME time blocks of consecutive channels, so the CAR column sums
    // of a time block are fetched from DRAM once and then hit in L2 for every other channel
    const int64_t ch = chan_fast ? blockIdx.x : blockIdx.y;
    const int64_t bpair = chan_fast ? blockIdx.y : blockIdx.x;
    const int U = kN - 2 * halo;
    const int64_t b0 = blk_begin + 2 * bpair, b1 = b0 + 1;
    const float* xr = x + ch * ldx;
    const float2* tw1 = tw;                               // [16][256]
    const float2* tw2 = tw + 16 * kHT;                    // [16][256]
    const float2* twAg = tw + 32 * kHT;                   // [16][256] W_256^{n1 k0} W_4096^{n0 k0}
    const int k0 = tid >> 4, k1 = tid & 15;
    const int nat0 = k0 + 17 * k1;                        // padi(k0 + 16 k1 + 256 k2) = nat0 + 272 k2

    twBs[k0 * kP18 + k1] = __ldg(&tw[48 * kHT + tid]);
    float2 v[16];
    {   // two real blocks as one complex signal, circular halo
        int64_t s0 = (b0 * U - halo) % T; if (s0 < 0) s0 += T;
        int64_t s1 = (b1 * U - halo) % T; if (s1 < 0) s1 += T;
        const bool has1 = b1 < blk_end;
        const bool wrap = (s0 + kN > T) || (s1 + kN > T);
        if (!wrap) {
#pragma unroll
            for (int n2 = 0; n2 < 16; ++n2) {
                const int i = 256 * n2 + tid;
                v[n2].x = xr[s0 + i];
                v[n2].y = has1 ? xr[s1 + i] : 0.f;
            }
            if (colsum) {     // common-average reference folded into the load: x - (1/n) sum_c w_c x_c
#pragma unroll
                for (int n2 = 0; n2 < 16; ++n2) {
                    const int i = 256 * n2 + tid;
                    v[n2].x = fmaf(-inv_count, __ldg(colsum + s0 + i), v[n2].x);
                    if (has1) v[n2].y = fmaf(-inv_count, __ldg(colsum + s1 + i), v[n2].y);
                }
            }
        } else {
#pragma unroll
            for (int n2 = 0; n2 < 16; ++n2) {
                const int i = 256 * n2 + tid;
                const int64_t i0 = (s0 + i) % T, i1 = (s1 + i) % T;
                v[n2].x = xr[i0];
                v[n2].y = has1 ? xr[i1] : 0.f;
                if (colsum) {
                    v[n2].x = fmaf(-inv_count, __ldg(colsum + i0), v[n2].x);
                    if (has1) v[n2].y = fmaf(-inv_count, __ldg(colsum + i1), v[n2].y);
                }
            }
        }
    }
    fft4096_pass1(v, bufA, tw1, tid);
    fft4096_finish(v, bufA, tw2, tid);
#pragma unroll
    for (int k2 = 0; k2 < 16; ++k2) bufB[nat0 + 272 * k2] = v[k2];      // natural-order spectrum Z[k]
    __syncthreads();
    // conjugate-symmetry split + gain + shift, stored conjugated (inverse = conj-forward FFT):
    //   block0: conj(Z[k] + conj(Z[N-k])) g,  block1: conj(-i (Z[k] - conj(Z[N-k]))) g,  k = shift + tid
    // bin tid = 16 n1 + n0 goes to [n0][n1] so that a pass-2 thread reads its 16 inputs with LDS.128
    const int sgi = k1 * kP18 + k0;
    for (int band = 0; band < nb; ++band) {
        const int k = shift.s[band] + tid;
        const float2 zk = bufB[padi(k)], zm = bufB[padi((kN - k) & (kN - 1))];
        const float gk = __ldg(&gain[band * 256 + tid]);
        SG[band * kSgBand + sgi] = make_float2((zk.x + zm.x) * gk, -(zk.y - zm.y) * gk);
        SG[(kFastBands + band) * kSgBand + sgi] = make_float2((zk.y + zm.y) * gk, (zk.x - zm.x) * gk);
    }
    float2 twA[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) twA[n1] = __ldg(&twAg[n1 * kHT + tid]);
    __syncthreads();

    const int n0 = k1;                                   // pass-2 role of this thread: (k0, n0)
    float* yr = y + ch * ldy;
    // The pass-2 -> pass-3 exchange of an inverse stays inside the 16 threads that share k0
    // (elements 288 k0 + 18 k1 + n0): one half-warp, so a warp-level barrier orders it and
    // every warp owns its slice of bufA.  The inter-pass twiddle W_256^{n0 k1} is applied on the
    // pass-3 side, inside the first butterfly layer.  bufB (free once SG is built) stages the
    // two output blocks in natural order, one half each: one CTA barrier per block.
    float2* p2 = bufA + 16 * kP18 * k0 + n0;
    const float4* p3 = reinterpret_cast<const float4*>(bufA + kP18 * tid);
    float2 wb[16];                                       // W_256^{n0 k1}, n0 = 0..15: registers for the whole CTA
#pragma unroll
    for (int i = 0; i < 16; ++i) wb[i] = twBs[kP18 * k1 + i];
    for (int sel = 0; sel < 2; ++sel) {
        float acc[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j] = 0.f;
        for (int band = 0; band < nb; ++band) {
            const float2* sgrow = SG + (sel * kFastBands + band) * kSgBand + n0 * kP18;
            switch (shift.nz[band]) {                    // uniform: bins >= 16 nz of this band's table are zero
                case 8: inverse_pass2<8>(v, sgrow, twA); break;
                case 9: inverse_pass2<9>(v, sgrow, twA); break;
                case 10: inverse_pass2<10>(v, sgrow, twA); break;
                case 11: inverse_pass2<11>(v, sgrow, twA); break;
                case 12: inverse_pass2<12>(v, sgrow, twA); break;
                case 13: inverse_pass2<13>(v, sgrow, twA); break;
                default: inverse_pass2<16>(v, sgrow, twA); break;
            }
            __syncwarp();                                // previous pass-3 reads of this slice are done
#pragma unroll
            for (int e = 0; e < 16; ++e) p2[kP18 * e] = v[e];
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float4 q = p3[i];
                v[2 * i] = make_float2(q.x, q.y);
                v[2 * i + 1] = make_float2(q.z, q.w);
            }
            dft16_tw<true, 16>(v, wb);
            // v[k2] = conj(z[t]) (times a unit phasor when shifted), t = k0 + 16 k1 + 256 k2
            if (ENV) {
#pragma unroll
                for (int j = EDGE ? 0 : 1; j < (EDGE ? 16 : 15); ++j) acc[j] += fast_sqrt(fmaf(v[j].x, v[j].x, v[j].y * v[j].y));
            } else {
#pragma unroll
                for (int j = EDGE ? 0 : 1; j < (EDGE ? 16 : 15); ++j) acc[j] += v[j].x;
            }
        }
        float* ob = reinterpret_cast<float*>(bufB) + sel * (kN + kN / 16);
#pragma unroll
        for (int k2 = EDGE ? 0 : 1; k2 < (EDGE ? 16 : 15); ++k2) ob[nat0 + 272 * k2] = acc[k2];
        __syncthreads();
        const int64_t bb = sel == 0 ? b0 : b1;
        if (bb < blk_end) {
            float* yo = yr + bb * U;
            const int64_t left = T - bb * U;
            const int lim = left < U ? (int)left : U;
            const float* os = ob + halo;
            for (int i = tid; i < lim; i += kHT) yo[i] = os[i + ((halo + i) >> 4)];     // padi(halo + i)
        }
    }
}

}  // namespace ecog

using namespace ecog;

extern "C" size_t ecog_hilbert_twiddle_floats(void) { return (size_t)(3 * 16 * kHT + 256) * 2; }

extern "C" int ecog_hilbert_twiddles(float* h_out) {
    if (!h_out) return fail(ECOG_E_VALUE, "ecog_hilbert_twiddles: null output");
    const double PI = 3.14159265358979323846;
    for (int e = 0; e < 16; ++e)
        for (int tid = 0; tid < kHT; ++tid) {
            // pass 1: W_256^{n1 k0}, n1 = tid >> 4, k0 = e
            double a1 = -2.0 * PI * (double)((tid >> 4) * e) / 256.0;
            h_out[2 * (e * kHT + tid) + 0] = (float)cos(a1);
            h_out[2 * (e * kHT + tid) + 1] = (float)sin(a1);
            // pass 2: W_4096^{n0 (k0 + 16 k1)}, k0 = tid >> 4, n0 = tid & 15, k1 = e
            double a2 = -2.0 * PI * (double)((tid & 15) * ((tid >> 4) + 16 * e)) / 4096.0;
            h_out[2 * (16 * kHT + e * kHT + tid) + 0] = (float)cos(a2);
            h_out[2 * (16 * kHT + e * kHT + tid) + 1] = (float)sin(a2);
            // fast path, pass-2 input: W_256^{n1 k0} W_4096^{n0 k0}, n1 = e, k0 = tid >> 4, n0 = tid & 15
            double a3 = -2.0 * PI * ((double)(e * (tid >> 4)) / 256.0 + (double)((tid & 15) * (tid >> 4)) / 4096.0);
            h_out[2 * (32 * kHT + e * kHT + tid) + 0] = (float)cos(a3);
            h_out[2 * (32 * kHT + e * kHT + tid) + 1] = (float)sin(a3);
        }
    for (int k1 = 0; k1 < 16; ++k1)
        for (int n0 = 0; n0 < 16; ++n0) {   // fast path, pass-2 output: W_256^{n0 k1}
            double a4 = -2.0 * PI * (double)(n0 * k1) / 256.0;
            h_out[2 * (48 * kHT + k1 * 16 + n0) + 0] = (float)cos(a4);
            h_out[2 * (48 * kHT + k1 * 16 + n0) + 1] = (float)sin(a4);
        }
    return ECOG_OK;
}

extern "C" int64_t ecog_hilbert_blocks(int64_t T, int32_t halo) {
    const int U = kN - 2 * halo;
    return U > 0 ? ceil_div(T, U) : 0;
}

extern "C" int ecog_hilbert_env(const float* d_x, float* d_y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                                const float* d_gain, int32_t nbands, int32_t rows, const int32_t* h_shift,
                                const int32_t* h_nz, int32_t halo, int32_t envelope, const float* d_twiddle,
                                const float* d_colsum, double inv_count, ecog_stream_t stream) {
    return ecog_hilbert_env_range(d_x, d_y, C, T, ldx, ldy, d_gain, nbands, rows, h_shift, h_nz, halo, envelope, d_twiddle,
                                  d_colsum, inv_count, 0, -1, stream);
}

extern "C" int ecog_hilbert_env_range(const float* d_x, float* d_y, int64_t C, int64_t T, int64_t ldx, int64_t ldy,
                                      const float* d_gain, int32_t nbands, int32_t rows, const int32_t* h_shift,
                                      const int32_t* h_nz, int32_t halo, int32_t envelope, const float* d_twiddle,
                                      const float* d_colsum, double inv_count, int64_t block_begin, int64_t block_end,
                                      ecog_stream_t stream) {
    if (C <= 0 || T <= 0 || ldx < T || ldy < T || C > 65535) return fail(ECOG_E_VALUE, "ecog_hilbert_env: bad shape");
    if (nbands < 1 || nbands > 64) return fail(ECOG_E_VALUE, "ecog_hilbert_env: 1..64 bands supported, got %d", nbands);
    if (rows != 1 && rows != 2 && rows != 4 && rows != 8)
        return fail(ECOG_E_VALUE, "ecog_hilbert_env: rows must be 1, 2, 4 or 8 (got %d)", rows);
    if (halo < 0 || 2 * halo >= kN / 2)
        return fail(ECOG_E_UNSUPPORTED, "ecog_hilbert_env: halo %d does not fit a %d-sample block", halo, kN);
    if (d_x == d_y) return fail(ECOG_E_VALUE, "ecog_hilbert_env: in-place operation is not supported");
    BandShift sh;
    for (int b = 0; b < 64; ++b) {
        sh.s[b] = (b < nbands && h_shift) ? h_shift[b] : 0;
        sh.nz[b] = (b < nbands && h_nz && rows == 1) ? h_nz[b] : 16;     // only the one-row fast path prunes
        if (sh.nz[b] < 1 || sh.nz[b] > 16)
            return fail(ECOG_E_VALUE, "ecog_hilbert_env: band %d: nz %d outside 1..16", b, sh.nz[b]);
        if (sh.nz[b] < 8) sh.nz[b] = 8;
        else if (sh.nz[b] > 13) sh.nz[b] = 16;
        if (sh.s[b] < 0 || sh.s[b] + rows * 256 > kHalf)
            return fail(ECOG_E_VALUE, "ecog_hilbert_env: band %d shift %d leaves the half spectrum", b, sh.s[b]);
        if (!envelope && sh.s[b] != 0)
            return fail(ECOG_E_VALUE, "ecog_hilbert_env: spectral shifts are only valid for the envelope");
    }
    const int U = kN - 2 * halo;
    const int64_t nBlocks = ceil_div(T, U);
    if (block_end < 0 || block_end > nBlocks) block_end = nBlocks;
    if (block_begin < 0 || block_begin % 2 || block_begin > block_end)
        return fail(ECOG_E_VALUE, "ecog_hilbert_env_range: block range [%lld, %lld) must start at an even block",
                    (long long)block_begin, (long long)block_end);
    if (block_begin == block_end) return ECOG_OK;
    dim3 grid((unsigned)ceil_div(block_end - block_begin, 2), (unsigned)C);
    const size_t smem = (size_t)2 * (kN + kN / 16) * sizeof(float2);
    const float2* tw = reinterpret_cast<const float2*>(d_twiddle);
    cudaStream_t st = (cudaStream_t)stream;
    if (rows == 1 && nbands <= kFastBands) {
        const size_t smem8 = kFastSmem;
        const int chan_fast = d_colsum != nullptr && grid.x <= 65535 ? 1 : 0;
        const dim3 grid8 = chan_fast ? dim3(grid.y, grid.x) : grid;
#define ECOG_HILBERT8(ENVV, EDGEV)                                                                                  \
    do {                                                                                                            \
        ECOG_TRY((smem_attr<hilbert_env8_kernel<ENVV, EDGEV>>(smem8)));                                             \
        hilbert_env8_kernel<ENVV, EDGEV><<<grid8, kHT, smem8, st>>>(d_x, d_y, T, ldx, ldy, d_gain, nbands, sh, halo, \
                                                                    tw, nBlocks, d_colsum, (float)inv_count,        \
                                                                    block_begin, block_end, chan_fast);             \
    } while (0)
        const bool edge = halo < 256;
        if (envelope && edge) ECOG_HILBERT8(true, true);
        else if (envelope) ECOG_HILBERT8(true, false);
        else if (edge) ECOG_HILBERT8(false, true);
        else ECOG_HILBERT8(false, false);
#undef ECOG_HILBERT8
        return check_launch("hilbert_env8");
    }
#define ECOG_HILBERT_LAUNCH(R)                                                                                   \
    do {                                                                                                         \
        ECOG_TRY((smem_attr<hilbert_env_kernel<R>>(smem)));                                                      \
        hilbert_env_kernel<R><<<grid, kHT, smem, st>>>(d_x, d_y, T, ldx, ldy, d_gain, nbands, sh, halo, envelope, \
                                                       tw, nBlocks, d_colsum, (float)inv_count, block_begin,     \
                                                       block_end);                                               \
    } while (0)
    if (rows == 1) ECOG_HILBERT_LAUNCH(1);
    else if (rows == 2) ECOG_HILBERT_LAUNCH(2);
    else if (rows == 4) ECOG_HILBERT_LAUNCH(4);
    else ECOG_HILBERT_LAUNCH(8);
#undef ECOG_HILBERT_LAUNCH
    return check_launch("hilbert_env");
}
