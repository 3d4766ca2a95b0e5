"""Oracle restatement of the reference's six preprocessing operators.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Every function takes a
``(C, T)`` array and returns a new array; ``fs`` is the sampling rate.  Citations
``ref:`` are relative to /root/reference, ``scipy:`` to scipy 1.18.1.
"""
from __future__ import annotations

import math
from typing import Iterable, List, Sequence, Tuple

import numpy as np
from scipy import fft as sp_fft
from scipy import linalg as sp_linalg
from scipy import signal as sp_signal


# ----------------------------------------------------------------------------
# Gaussian "Hilbert" filter bank      ref: preprocess/signal/frequency_filter.py:80-184
# ----------------------------------------------------------------------------
def normalise_freq_ranges(freq_ranges) -> List[Tuple[float, float]]:
    """ref: frequency_filter.py:121-124.  A tuple or a flat list starting with a
    float is ONE range.  (A flat list of two ints raises TypeError in the
    reference, Appendix B2; the oracle accepts ints like the product does.)"""
    if isinstance(freq_ranges, tuple):
        return [tuple(freq_ranges)]
    if isinstance(freq_ranges[0], (int, float, np.integer, np.floating)):
        return [tuple(freq_ranges)]
    return [tuple(r) for r in freq_ranges]


def gaussian_bank(freq_ranges, f0=0.018, octspace=1.0 / 7.0,
                  filterbank_bias=math.log10(0.39), filterbank_slope=0.5):
    """Centre frequencies and Gaussian widths, ref: frequency_filter.py:128-152.
    The centre frequencies come from the *iterated product* ``f *= 2**octspace``
    (not a closed form) so the last bits match the reference."""
    cfs, sds = [], []
    for rng in normalise_freq_ranges(freq_ranges):
        if len(rng) != 2:
            raise ValueError("Each frequency range must be a (min_freq, max_freq) pair.")
        lo, hi = rng
        max_oct = math.log2(hi / f0)
        f = f0
        while math.log2(f / f0) < max_oct:
            if f >= lo:
                cfs.append(f)
                sds.append(10 ** (filterbank_bias + filterbank_slope * math.log10(f)))
            f = f * (2 ** octspace)
    return np.array(cfs), np.array(sds) * np.sqrt(2)


def analytic_mask(T: int) -> np.ndarray:
    """1 / 2 ... 2 / 1 / 0 ... 0 multiplier, ref: frequency_filter.py:158-165."""
    m = np.zeros(T)
    m[0] = 1
    if T % 2 == 0:
        m[1:T // 2] = 2
        m[T // 2] = 1
    else:
        m[1:(T + 1) // 2] = 2
    return m


def hilbert_filter(data, fs, freq_ranges, f0=0.018, octspace=1.0 / 7.0,
                   filterbank_bias=math.log10(0.39), filterbank_slope=0.5,
                   envelope=True) -> np.ndarray:
    """ref: frequency_filter.py:154-184.  Whole-record FFT, one Gaussian x analytic
    mask per band, inverse FFT, |.| (or real part), mean over bands.  float64 out.
    Looped per channel so a 60-min row does not need the reference's (C,T,bands)
    temporary (frequency_filter.py:170); the arithmetic per element is the same."""
    data = np.asarray(data)
    C, T = data.shape
    cfs, sds = gaussian_bank(freq_ranges, f0, octspace, filterbank_bias, filterbank_slope)
    freqs = np.fft.fftfreq(T, d=1.0 / fs)
    mask = analytic_mask(T)
    kernels = []
    for fc, sd in zip(cfs, sds):
        H = np.exp(-0.5 * ((freqs - fc) / sd) ** 2)
        H[0] = 0
        kernels.append(H * mask)
    out = np.zeros((C, T))
    # one batched transform like the reference (:167): float32 input gives a
    # complex64 spectrum, and pocketfft's row batching decides its last bits
    spectrum = sp_fft.fft(data, axis=1)
    for ch in range(C):
        X = spectrum[ch]
        bands = np.empty((T, len(kernels)))
        for i, K in enumerate(kernels):
            z = sp_fft.ifft(X * K)
            bands[:, i] = np.abs(z) if envelope else z.real
        out[ch] = bands.mean(axis=1)
    return out


# ----------------------------------------------------------------------------
# Butterworth zero-phase / causal    ref: preprocess/signal/frequency_filter.py:187-229
# ----------------------------------------------------------------------------
def lfilter_zi(b: np.ndarray, a: np.ndarray) -> np.ndarray:
    """Steady-state DF2T state for a unit step, scipy 1.18.1: signal/_signaltools.py:4330
    ``lfilter_zi``: ``y_inf = sum(b)/sum(a)``; ``zi[k] = sum_{j>k} (b[j] - y_inf a[j])``
    (reverse cumulative sum).  NOTE scipy 1.11.4 (the reference's pin) solved
    ``(I - companion(a)^T) zi = b[1:] - a[1:] b[0]`` instead; for a 4 Hz notch at
    2 kHz the two differ by 3e-7, which moves the filtfilt edges by ~2e-5 -- the
    container's scipy is the oracle of record."""
    b = np.atleast_1d(b)
    a = np.atleast_1d(a)
    if a[0] != 1:
        b, a = b / a[0], a / a[0]
    y_inf = b.sum() / a.sum()
    n = max(len(a), len(b))
    a = np.pad(a, (0, n - len(a)))
    b = np.pad(b, (0, n - len(b)))
    return np.flip(np.cumsum(np.flip(b - y_inf * a)))[1:]


def odd_extension(x: np.ndarray, n: int) -> np.ndarray:
    """Point-reflect ``n`` samples about each end (scipy ``odd_ext``)."""
    left = 2 * x[..., :1] - x[..., n:0:-1]
    right = 2 * x[..., -1:] - x[..., -2:-(n + 2):-1]
    return np.concatenate((left, x, right), axis=-1)


def filtfilt_pad(b, a, x, dtype=np.float64, reference_edges: bool = False) -> np.ndarray:
    """scipy ``filtfilt`` defaults (padtype='odd', padlen=3*max(len(a),len(b)),
    method='pad'), scipy: signal/_signaltools.py:4892-4960, as called by
    ref: frequency_filter.py:226-227.  ``dtype=np.longdouble`` gives the
    extended-precision "truth" of SURVEY.md section 8c.  ``reference_edges=True`` keeps the two
    edge ingredients exactly as the float64 reference forms them -- the odd extension in the INPUT's
    dtype and ``lfilter_zi`` in float64 -- and widens only the recursion: the error that is left is the
    round-off of the 8th-order direct form, which is what the long-double rule is about."""
    b64 = np.asarray(b, dtype=np.float64)
    a64 = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=dtype)
    a = np.asarray(a, dtype=dtype)
    x = np.asarray(x)
    if dtype != np.float64 and not reference_edges:
        x = x.astype(dtype)
    edge = 3 * max(len(a), len(b))
    if x.shape[-1] <= edge:
        raise ValueError(
            f"The length of the input vector x must be greater than padlen, which is {edge}.")
    ext = odd_extension(x, edge)     # in the INPUT's dtype (float32 data -> float32 pad), as scipy does
    if reference_edges:
        ext = ext.astype(dtype)
        zi = lfilter_zi(b64, a64).astype(dtype)
    else:
        zi = lfilter_zi(b, a)        # same arithmetic width as the recursion
    zshape = [1] * x.ndim
    zshape[-1] = zi.size
    zi = zi.reshape(zshape)
    y, _ = sp_signal.lfilter(b, a, ext, axis=-1, zi=zi * ext[..., :1])
    y, _ = sp_signal.lfilter(b, a, y[..., ::-1], axis=-1, zi=zi * y[..., -1:])
    y = y[..., ::-1]
    return y[..., edge:-edge]


def butter_filter(data, freqs, fs, order=4, causal=False, filter_type="bandpass"):
    """ref: frequency_filter.py:218-229."""
    wn = np.asarray(freqs, dtype=float) / (0.5 * fs)
    if causal:
        sos = sp_signal.butter(order, wn, btype=filter_type, output="sos")
        return sp_signal.sosfilt(sos, np.asarray(data), axis=-1)
    b, a = sp_signal.butter(order, wn, btype=filter_type)
    return filtfilt_pad(b, a, np.asarray(data))


def fir_bandpass_filter(data, fs, order, center_frequencies):
    """ref: frequency_filter.py:260-274.  Note the reference hands firwin the
    *Nyquist-normalised* edges together with ``fs=fs`` (:265-268) -- reproduced
    as is; the accumulator has the input's dtype (``zeros_like``, :261)."""
    data = np.asarray(data)
    acc = np.zeros_like(data)
    nyq = 0.5 * fs
    for fc in center_frequencies:
        lo, hi = fc * 0.9 / nyq, fc * 1.1 / nyq
        taps = sp_signal.firwin(order + 1, [lo, hi], pass_zero=False, fs=fs)
        acc += sp_signal.lfilter(taps, 1.0, data, axis=-1).astype(acc.dtype, copy=False)
    acc /= len(center_frequencies)
    return acc


def frequency_filter(data, fs, bands: Sequence[dict]):
    """ref: frequency_filter.py:35-77.  Every band runs on the SAME input and the
    results are concatenated along the channel axis."""
    if bands is None:
        raise ValueError("bands must be specified in params.")
    outs = []
    for cfg in bands:
        method = cfg.get("method", "hilbert")
        p = cfg.get("params", {}) or {}
        if method == "hilbert":
            if "freq_ranges" not in p:
                raise ValueError("Hilbert filter requires 'freq_ranges' in params.")
            outs.append(hilbert_filter(data, fs, **p))
        elif method == "butter":
            if "freqs" not in p:
                raise ValueError("Butterworth filter requires 'freq_range' in params.")
            outs.append(butter_filter(data, fs=fs, **p))
        elif method == "fir":
            if "order" not in p or "center_frequencies" not in p:
                raise ValueError("FIR filter requires 'order' and 'center_frequencies' in params.")
            outs.append(fir_bandpass_filter(data, fs, p["order"], p["center_frequencies"]))
    return np.concatenate(outs, axis=0)


# ----------------------------------------------------------------------------
# CAR / z-scores
# ----------------------------------------------------------------------------
def car_rereference(data, exclude_channels=()):
    """ref: preprocess/signal/car_rereference.py:26-39."""
    data = np.asarray(data)
    if not isinstance(exclude_channels, (list, tuple)):
        raise ValueError("exclude_channels must be a list of integers.")
    if any(ch < 0 or ch >= data.shape[0] for ch in exclude_channels):
        raise ValueError("exclude_channels contains invalid channel indices.")
    keep = np.ones(data.shape[0], dtype=bool)
    keep[list(exclude_channels)] = False
    return data - np.mean(data[keep, :], axis=0, keepdims=True)


def channel_zscore(data, preserve_nans=True):
    """ref: preprocess/signal/channel_zscore.py:20-29 (population std)."""
    data = np.asarray(data)
    with np.errstate(divide="ignore", invalid="ignore"):
        z = (data - data.mean(axis=1, keepdims=True)) / data.std(axis=1, keepdims=True)
    if not preserve_nans:
        z[np.isnan(z)] = 0
    return z


def zscore_rereference(data, fs, interval):
    """ref: preprocess/signal/zscore_rereference.py:24-28,52-70."""
    data = np.asarray(data)
    s, e = interval
    s, e = int(s * fs), int(e * fs)
    if s < 0 or e > data.shape[1]:
        raise ValueError("Reference time indices are out of bounds.")
    if s >= e:
        raise ValueError("Start time must be less than end time.")
    ref = data[:, s:e]
    with np.errstate(divide="ignore", invalid="ignore"):
        return (data - ref.mean(axis=1, keepdims=True)) / ref.std(axis=1, keepdims=True)


def rolling_zscore(data, fs, window_length=10, preserve_nans=True):
    """ref: preprocess/signal/rolling_zscore.py:28-49.  pandas trailing window of
    ``int(window_length*fs)`` samples, ``min_periods=1``, mean and SAMPLE std
    (ddof=1), so sample 0 is 0/NaN = NaN.  Restated with float64 prefix sums of
    the per-row mean-shifted data (pandas uses an add/remove online update; both
    agree to ~1e-12).  float64 out."""
    data = np.asarray(data, dtype=np.float64)
    w = int(window_length * fs)
    if w <= 1:
        raise ValueError("window_size must be greater than 1.")
    C, T = data.shape
    shift = data.mean(axis=1, keepdims=True)
    d = (data - shift).astype(np.longdouble)
    c1 = np.concatenate([np.zeros((C, 1), np.longdouble), np.cumsum(d, axis=1)], axis=1)
    c2 = np.concatenate([np.zeros((C, 1), np.longdouble), np.cumsum(d * d, axis=1)], axis=1)
    t = np.arange(T)
    lo = np.maximum(t + 1 - w, 0)
    n = (t + 1 - lo).astype(np.longdouble)
    s1 = c1[:, t + 1] - c1[:, lo]
    s2 = c2[:, t + 1] - c2[:, lo]
    mean = s1 / n
    with np.errstate(divide="ignore", invalid="ignore"):
        var = (s2 - s1 * s1 / n) / (n - 1)
        var = np.where(n > 1, np.maximum(var, 0), np.nan)
        z = (d - mean) / np.sqrt(var)
    z = z.astype(np.float64)
    if not preserve_nans:
        z[np.isnan(z)] = 0
    return z


# ----------------------------------------------------------------------------
# FFT resample                       ref: preprocess/signal/downsample.py:21-27
# ----------------------------------------------------------------------------
def resample_length(T: int, fs, target) -> int:
    """ref: downsample.py:23-24 (float64 ratio, truncation)."""
    return int(T * (target / fs))


def fft_resample(data, num: int) -> np.ndarray:
    """scipy ``resample`` for real input, no window, scipy: signal/_signaltools.py
    (1.18.1) ``resample``: rfft, keep ``m//2+1`` bins, fix the unpaired bin when
    ``m`` is even and the length changes, irfft to ``num`` samples scaled by
    ``num/T``.  dtype preserved (float32 in -> float32 out, pocketfft single)."""
    data = np.asarray(data)
    T = data.shape[-1]
    m = min(num, T)
    X = sp_fft.rfft(data, axis=-1)[..., : m // 2 + 1].copy()
    if m % 2 == 0 and num != T:
        X[..., m // 2] *= 2 if num < T else 0.5
    return sp_fft.irfft(X / (T / num), n=num, axis=-1)


def downsample(data, fs, target=400):
    """ref: downsample.py:19-29; returns (array, new signal_freq)."""
    data = np.asarray(data)
    return fft_resample(data, resample_length(data.shape[1], fs, target)), target
