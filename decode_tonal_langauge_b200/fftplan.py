"""Host-side plans and tables for the whole-row FFT resampler (K5) and the block
Hilbert kernel (K4).  Float64 everywhere, rounded to float32 once.

Resample of a real row of T samples to ``num`` samples (ref: preprocess/signal/downsample.py:21-27,
scipy.signal.resample): the device computes

  forward : z[j] = x[2j] + i x[2j+1] (N = T/2 complex points), Z = FFT_N(z) as a
            four-step transform N = n_a * n_b: column pass (length n_a, stride n_b,
            times W_N^{q c}) stored transposed, then a second column pass (length n_b);
  repack  : rfft bins X[k] from Z[k], Z[N-k]; truncate / scale / Nyquist rule; fold
            the half spectrum into G[k] (N' = num/2 complex points);
  inverse : g = IFFT_{N'}(G) run as a conjugated forward four-step transform;
            y[2j] = Re g[j], y[2j+1] = Im g[j].
"""
from __future__ import annotations

import functools
from dataclasses import dataclass
from typing import List, Tuple

import numpy as np

MAX_AXIS = 2800          # longest in-shared-memory FFT: n * (8 + 1) * 8 B <= ~200 KB
MAX_AXIS_NARROW = 5600   # with 4-column tiles (n * 5 * 8 B): only the chirp-z path goes this far
TILE_W = 8
BIG_SPLIT = 4096         # two-level table for the four-step twiddles W_N^e, e = hi*4096 + lo


def factorize(n: int) -> List[int]:
    """Radix sequence over {4, 2, 3, 5}; raises for other prime factors."""
    radices = []
    m = n
    while m % 4 == 0:
        radices.append(4)
        m //= 4
    for r in (2, 3, 5):
        while m % r == 0:
            radices.append(r)
            m //= r
    if m != 1:
        raise NotImplementedError(
            f"FFT length {n} has a prime factor other than 2, 3, 5 (left {m}); "
            "non-smooth lengths need Bluestein (SURVEY.md section 8f row f3)")
    return radices


def digit_reversal(n: int, radices: List[int]) -> np.ndarray:
    """perm[i] = shared-memory row where input sample i must be placed so that the
    in-place decimation-in-time stages (radices[0] first) leave natural order."""
    def rec(idx: np.ndarray, length: int, rs: List[int]) -> np.ndarray:
        if not rs:
            return np.zeros_like(idx)
        r = rs[-1]
        return (idx % r) * (length // r) + rec(idx // r, length // r, rs[:-1])
    return rec(np.arange(n, dtype=np.int64), n, list(radices)).astype(np.int32)


def roots(n: int, sign: float = -1.0) -> np.ndarray:
    """(n, 2) float32 table of W_n^k = exp(sign * 2 pi i k / n)."""
    k = np.arange(n, dtype=np.float64)
    ang = sign * 2.0 * np.pi * k / n
    return np.stack([np.cos(ang), np.sin(ang)], axis=1).astype(np.float32)


def split_size(N: int, max_axis: int = MAX_AXIS) -> Tuple[int, int]:
    """N = n_a * n_b with both factors 2-3-5 smooth, <= max_axis, n_b a multiple of TILE_W
    where possible and the pair as balanced as possible.  (n_a, 1) for short rows."""
    factorize(N)
    if N <= MAX_AXIS:
        return N, 1
    best = None
    for na in range(2, max_axis + 1):
        if N % na:
            continue
        nb = N // na
        if nb > max_axis:
            continue
        score = (0 if nb % TILE_W == 0 else 1, 0 if na % TILE_W == 0 else 1, abs(na - nb))
        if best is None or score < best[0]:
            best = (score, na, nb)
    if best is None:
        raise NotImplementedError(f"FFT length {N} does not split into two factors <= {max_axis}")
    return best[1], best[2]


@dataclass(frozen=True)
class AxisPlan:
    n: int
    radices: Tuple[int, ...]
    perm: np.ndarray        # int32 (n,)
    tw: np.ndarray          # float32 (n, 2)


def axis_plan(n: int) -> AxisPlan:
    rs = factorize(n) if n > 1 else []
    return AxisPlan(n, tuple(rs), digit_reversal(n, rs), roots(max(n, 1)))


@dataclass(frozen=True)
class BigPlan:
    N: int
    a: AxisPlan             # first (column) pass, length n_a, stride n_b
    b: AxisPlan             # second pass, length n_b
    tw_hi: np.ndarray       # float32 (ceil(N / BIG_SPLIT), 2): W_N^{hi * BIG_SPLIT}
    tw_lo: np.ndarray       # float32 (BIG_SPLIT, 2):          W_N^{lo}
    tw_q: np.ndarray        # float64 (n_a, 2):                W_N^q (base of the column powers)


def big_plan(N: int, max_axis: int = MAX_AXIS) -> BigPlan:
    na, nb = split_size(N, max_axis)
    hi = np.arange(-(-N // BIG_SPLIT), dtype=np.float64) * BIG_SPLIT
    lo = np.arange(BIG_SPLIT, dtype=np.float64)
    f = lambda e: np.stack([np.cos(-2 * np.pi * e / N), np.sin(-2 * np.pi * e / N)], axis=1).astype(np.float32)
    q = np.arange(na, dtype=np.float64)
    tw_q = np.stack([np.cos(-2 * np.pi * q / N), np.sin(-2 * np.pi * q / N)], axis=1)
    return BigPlan(N, axis_plan(na), axis_plan(nb), f(hi), f(lo), np.ascontiguousarray(tw_q))


@dataclass(frozen=True)
class ResamplePlan:
    T: int
    num: int
    fwd: BigPlan
    inv: BigPlan
    tw_T: np.ndarray        # float32 (num/2 + 1, 2): W_T^k        (rfft untangle)
    tw_num: np.ndarray      # float32 (num/2, 2):     W_num^{-k}   (irfft fold)


@functools.lru_cache(maxsize=16)
def resample_plan(T: int, num: int) -> ResamplePlan:
    if T % 2 or num % 2:
        raise NotImplementedError(
            f"FFT resample {T} -> {num}: odd lengths are not implemented yet (SURVEY.md 8f row f3)")
    if num < 2 or T < 2:
        raise ValueError("resample needs at least two samples in and out")
    Nh = num // 2
    k = np.arange(Nh + 1, dtype=np.float64)
    tw_T = np.stack([np.cos(-2 * np.pi * k / T), np.sin(-2 * np.pi * k / T)], axis=1).astype(np.float32)
    k = k[:Nh]
    tw_num = np.stack([np.cos(2 * np.pi * k / num), np.sin(2 * np.pi * k / num)], axis=1).astype(np.float32)
    return ResamplePlan(T, num, big_plan(T // 2), big_plan(Nh), tw_T, tw_num)


# ------------------------------------------------------------ chirp-z resampling
def next_smooth_even(n: int) -> int:
    """Smallest even 2-3-5 smooth integer >= n that splits into two shared-memory axes."""
    n = max(int(n), 2)
    while True:
        m = n
        for p in (2, 3, 5):
            while m % p == 0:
                m //= p
        if m == 1 and n % 2 == 0:
            try:
                split_size(n, MAX_AXIS_NARROW)
                return n
            except NotImplementedError:
                pass
        n += 1


def _chirp(n: int, L: int) -> np.ndarray:
    """exp(+i pi k^2 / L), k < n, with the phase reduced exactly (integers) modulo 2 L."""
    k = np.arange(n, dtype=np.int64)
    ph = (k * k) % (2 * L)
    return np.exp(1j * np.pi * ph.astype(np.float64) / L)


def _c2(a: np.ndarray) -> np.ndarray:
    """complex128 -> (n, 2) float32"""
    return np.ascontiguousarray(np.stack([a.real, a.imag], axis=1).astype(np.float32))


@dataclass(frozen=True)
class CztPlan:
    """Bluestein realisation of scipy.signal.resample(x, num) for ANY row length T
    (ref: preprocess/signal/downsample.py:21-27; real TDT rates give non-smooth, odd lengths).

      forward  X[k] = conj(w[k]) * sum_n (x[n] conj(w[n])) w[k - n],   w[n] = exp(i pi n^2 / T),
               k < K = min(num, T)//2 + 1, as a circular convolution of smooth length M1 >= T + K - 1;
      inverse  y[m] = Re( v[m] * sum_k (G[k] v[k]) conj(v[m - k]) ),   v[j] = exp(i pi j^2 / num),
               G[k] = X[k] * (num/T) * Nyquist rule * hermitian weight / num, length M2 >= K + num - 1.
    All tables float64 -> float32 once; FB1 / FB2 are the FFTs of the chirp kernels (1/M folded in)."""
    T: int
    num: int
    K: int
    M1: int
    M2: int
    fft1: BigPlan
    fft2: BigPlan
    pre: np.ndarray      # (T, 2)   conj(w[n])
    FB1: np.ndarray      # (M1, 2)
    mid: np.ndarray      # (K, 2)   conj(w[k]) * scale[k] * h[k] / num * v[k]  (* bin_gain[k])
    FB2: np.ndarray      # (M2, 2)
    post: np.ndarray     # (num, 2) v[m]


@functools.lru_cache(maxsize=8)
def _czt_plan(T: int, num: int, gain_key) -> CztPlan:
    from scipy import fft as sp_fft
    if T < 2 or num < 2:
        raise ValueError("resample needs at least two samples in and out")
    m = min(num, T)
    K = m // 2 + 1
    M1 = next_smooth_even(T + K - 1)
    M2 = next_smooth_even(K + num - 1)
    w = _chirp(max(T, K), T)
    b = np.zeros(M1, dtype=np.complex128)
    b[:K] = w[:K]
    b[M1 - np.arange(1, T)] = w[1:T]
    FB1 = sp_fft.fft(b) / M1
    scale = np.full(K, num / T)
    if m % 2 == 0 and num != T:                     # scipy's unpaired-bin rule
        scale[m // 2] *= 2.0 if num < T else 0.5
    h = np.full(K, 2.0)
    h[0] = 1.0
    if num % 2 == 0 and K - 1 == num // 2:
        h[K - 1] = 1.0                              # the output's own Nyquist bin is not doubled
    v = _chirp(max(num, K), num)
    mid = np.conj(w[:K]) * scale * h / num * v[:K]
    b2 = np.zeros(M2, dtype=np.complex128)
    b2[:num] = np.conj(v[:num])
    b2[M2 - np.arange(1, K)] = np.conj(v[1:K])
    FB2 = sp_fft.fft(b2) / M2
    return CztPlan(T, num, K, M1, M2, big_plan(M1, MAX_AXIS_NARROW), big_plan(M2, MAX_AXIS_NARROW),
                   _c2(np.conj(w[:T])), _c2(FB1), mid, _c2(FB2), _c2(v[:num]))


def czt_plan(T: int, num: int) -> CztPlan:
    return _czt_plan(int(T), int(num), None)


@dataclass(frozen=True)
class BluesteinPlan:
    """DFT of ANY length T as a circular convolution of smooth length M >= 2 T - 1 (chirp-z with all T bins):

      forward  X[k] = conj(w[k]) * sum_n (x[n] conj(w[n])) w[k - n]
      inverse  x[n] = (1/T) w[n] * sum_k (X[k] w[k]) conj(w[n - k]),      w[j] = exp(i pi j^2 / T)

    Used by the whole-record Gaussian-Hilbert path (ref: frequency_filter.py:154-184) when the row length has a
    prime factor other than 2, 3, 5 (real TDT rates).  FBf / FBi: FFTs of the two chirp kernels, 1/M folded in."""
    T: int
    M: int
    fft: BigPlan
    w: np.ndarray        # (T,) complex128 chirp
    FBf: np.ndarray      # (M, 2) float32
    FBi: np.ndarray      # (M, 2) float32


@functools.lru_cache(maxsize=4)
def bluestein_plan(T: int) -> BluesteinPlan:
    from scipy import fft as sp_fft
    T = int(T)
    M = next_smooth_even(2 * T - 1)
    w = _chirp(T, T)
    b = np.zeros(M, dtype=np.complex128)
    b[:T] = w
    b[M - np.arange(1, T)] = w[1:T]
    return BluesteinPlan(T, M, big_plan(M, MAX_AXIS_NARROW), w, _c2(sp_fft.fft(b) / M), _c2(sp_fft.fft(np.conj(b)) / M))


# ------------------------------------------------------- two-stage resampling
FIR_MAX_TAPS = 256
FIR_ATTENUATION_DB = 120.0


@dataclass(frozen=True)
class PreDecimation:
    D: int                  # integer decimation factor of the FIR stage
    taps: np.ndarray        # float32 (ntaps,), ntaps % 4 == 0 (zero padded), DC gain 1
    offset: int             # y1[m] = sum_j taps[j] x[(m D + j - offset) mod T]; offset % 4 == 0
    bin_gain: np.ndarray    # float32 (num/2 + 1,): 1 / H1[k], H1 = DFT_T of the (float32) taps
    halfband: tuple = None  # D == 4 as two half-band stages: (stage1, stage2), each float32 (1 + K,) = centre tap, odd taps


HALFBAND_K = (8, 21)        # odd tap pairs the kernel is built for (csrc/firdecim.cu: kHbK1, kHbK2)


def _halfband_stage(f_pass: float, kmax: int):
    """Half-band low-pass (cut-off at half the Nyquist frequency, every second tap zero) that stops [1 - f_pass, 1]
    by FIR_ATTENUATION_DB: the shortest Kaiser design with at most ``kmax`` odd tap pairs, float32; None if none does.
    Returns (centre tap, odd taps g_i at distance 2 i + 1) as one float32 vector."""
    from scipy import signal as sp_signal
    f = np.linspace(1.0 - f_pass, 1.0, 4001)
    for K in range(2, kmax + 1):
        for att in (FIR_ATTENUATION_DB, FIR_ATTENUATION_DB + 4.0, FIR_ATTENUATION_DB + 8.0):
            n = 4 * K - 1
            h = sp_signal.firwin(n, 0.5, window=("kaiser", sp_signal.kaiser_beta(att)))
            c = 2 * K - 1
            st = np.concatenate([[h[c]], h[c + 1::2]]).astype(np.float32)
            i = np.arange(K, dtype=np.float64)
            H = float(st[0]) + 2.0 * np.sum(st[1:].astype(np.float64)[None, :] * np.cos(np.pi * f[:, None] * (2 * i + 1)[None, :]), axis=1)
            if 20.0 * np.log10(np.max(np.abs(H)) + 1e-300) <= -FIR_ATTENUATION_DB:
                return st
    return None


def _halfband_response(st: np.ndarray, k: np.ndarray, period: int) -> np.ndarray:
    """Exact response of a (float32) half-band stage at the bins k of a `period`-sample row (zero phase)."""
    i = np.arange(st.size - 1, dtype=np.float64)
    return float(st[0]) + 2.0 * np.sum(st[1:].astype(np.float64)[None, :]
                                         * np.cos(2.0 * np.pi * k[:, None] * (2 * i + 1)[None, :] / period), axis=1)


@functools.lru_cache(maxsize=16)
def predecimation(T: int, num: int):
    """FIR pre-decimator for ``resample(T -> num)`` or None when the ratio is too small.

    The brick wall keeps bins k <= num/2 of the T-point spectrum.  Decimating by D folds bin
    T/D - k onto k, so the FIR must pass [0, num/2] (any non-zero gain: it is divided out
    bin by bin) and stop [T/D - num/2, T/2] by FIR_ATTENUATION_DB.  Kaiser design in float64,
    rounded to float32 once; the compensation uses the ROUNDED taps' exact response."""
    from scipy import signal as sp_signal
    from scipy import fft as sp_fft
    if num >= T:
        return None
    if T % 4 == 0 and T // 4 >= 1.15 * num and T > 4 * (4 * HALFBAND_K[1] + 64):
        # decimation by 4 as two half-band stages (19 instead of 40 products per sample at 2 kHz -> 400 Hz)
        s1 = _halfband_stage(num / T, HALFBAND_K[0])
        s2 = _halfband_stage(2.0 * num / T, HALFBAND_K[1])
        if s1 is not None and s2 is not None:
            k = np.arange(num // 2 + 1, dtype=np.float64)
            H = _halfband_response(s1, k, T) * _halfband_response(s2, k, T // 2)
            if np.min(H) >= 0.5:
                return PreDecimation(4, np.zeros(4, dtype=np.float32), 0, (1.0 / H).astype(np.float32), (s1, s2))
    for D in (4, 2):
        if T % D:
            continue
        T1 = T // D
        if T1 < 1.15 * num:
            continue
        f_pass = num / T                      # in units of the input Nyquist frequency
        f_stop = 2.0 / D - num / T
        ntaps, beta = sp_signal.kaiserord(FIR_ATTENUATION_DB, f_stop - f_pass)
        ntaps |= 1                            # odd -> symmetric about an integer sample
        if ntaps + 3 > FIR_MAX_TAPS or ntaps >= T:
            continue
        h = sp_signal.firwin(ntaps, 0.5 * (f_pass + f_stop), window=("kaiser", beta))
        c = (ntaps - 1) // 2
        lead = (-c) % 4                       # leading zeros so that the centre offset is a multiple of 4
        total = -(-(lead + ntaps) // 4) * 4
        taps = np.zeros(total, dtype=np.float32)
        taps[lead:lead + ntaps] = h.astype(np.float32)
        offset = c + lead
        # exact response of the rounded taps on the T-point grid (zero-phase arrangement)
        k = np.zeros(T, dtype=np.float64)
        idx = (np.arange(total) - offset) % T
        np.add.at(k, idx, taps.astype(np.float64))
        H1 = sp_fft.rfft(k)[: num // 2 + 1]
        if np.max(np.abs(H1.imag)) > 1e-9 or np.min(H1.real) < 0.5:
            continue                          # not a usable pass band (cannot happen for a Kaiser low-pass)
        return PreDecimation(D, taps, int(offset), (1.0 / H1.real).astype(np.float32))
    return None


# --------------------------------------------------------------------- Hilbert
HILBERT_N = 4096


def hilbert_gain(cfs: np.ndarray, sds: np.ndarray, fs: float, envelope: bool = True):
    """Per-band gain tables on the 4096-point block grid.

    Returns (gain (nb, rows*256) float32, shift (nb,) int32, rows).  gain = Gaussian x analytic
    factor 2, times 1/2 (conjugate-symmetry split), 1/N (inverse FFT) and 1/nb (mean over
    bands); ref: frequency_filter.py:158-175,184 -- H[0] = 0.  Bins whose gain is below 1e-9
    of the band peak are dropped (6.4 sigma; the dropped tails carry < 1e-9 of a line's amplitude,
    two orders below the float32 arithmetic); for the envelope each band is shifted down to its
    first kept bin (|z| is invariant to a spectral shift) so that it fits rows*256 bins and the
    kernel can skip the zero tail of the window (hilbert_nz)."""
    N = HILBERT_N
    f = np.arange(N // 2, dtype=np.float64) * fs / N
    g = np.exp(-0.5 * ((f[None, :] - cfs[:, None]) / sds[:, None]) ** 2)
    g[:, 0] = 0.0
    keep = g >= 1e-9 * g.max(axis=1, keepdims=True)
    lo = np.array([np.argmax(k) for k in keep])
    hi = np.array([len(k) - np.argmax(k[::-1]) for k in keep])       # one past the last kept bin
    if not envelope:
        lo[:] = 0
    width = int((hi - lo).max())
    rows = next(r for r in (1, 2, 4, 8) if width <= r * 256)
    lo = np.minimum(lo, N // 2 - rows * 256)
    out = np.zeros((len(cfs), rows * 256), dtype=np.float64)
    for b in range(len(cfs)):
        seg = g[b, lo[b]:lo[b] + rows * 256].copy()
        seg[np.arange(lo[b], lo[b] + rows * 256) >= hi[b]] = 0.0
        out[b] = seg
    out *= 2.0 * 0.5 / N / len(cfs)
    return np.ascontiguousarray(out, dtype=np.float32), lo.astype(np.int32), rows


def hilbert_nz(gain: np.ndarray) -> np.ndarray:
    """Per band: number of leading 16-bin groups of the gain table that hold a non-zero entry
    (the `h_nz` promise of ecog_hilbert_env; at least 1)."""
    nzb = np.array([(np.flatnonzero(g)[-1] + 1 if np.any(g) else 1) for g in gain])
    return np.maximum(-(-nzb // 16), 1).astype(np.int32)


def hilbert_halo(cfs: np.ndarray, sds: np.ndarray, fs: float, T: int, nsigma: float = 5.0) -> int:
    """Samples of context each block needs: nsigma standard deviations of the widest time-domain
    Gaussian (5 sigma leaves erfc(5/sqrt 2) = 5.7e-7 of the kernel weight outside the halo, the
    worst-case relative error; measured 1.6e-7 against the whole-record reference, the same as
    with 6.5 sigma where the float32 tables dominate; 4.5 sigma measures 4.5e-6).  Raises NotImplementedError when the bank cannot
    be evaluated block-wise (kernels too long, or gain not negligible at DC / Nyquist)."""
    sigma_t = 1.0 / (2.0 * np.pi * sds)                 # seconds, per band
    halo = int(np.ceil(nsigma * sigma_t.max() * fs))
    edge = np.maximum(np.exp(-0.5 * (cfs / sds) ** 2), np.exp(-0.5 * ((fs / 2 - cfs) / sds) ** 2))
    if halo >= HILBERT_N // 4 or edge.max() > 1e-12:
        raise NotImplementedError(
            "this Gaussian bank needs the whole-record FFT path (low-frequency or near-Nyquist "
            f"bands: halo {halo} samples, edge gain {edge.max():.1e}); only the block-wise "
            "kernel is implemented (SURVEY.md section 8f row f3)")
    return max(halo, 1)
