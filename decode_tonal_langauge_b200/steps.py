"""The reference's six preprocessing operators with its plug-in contract

    run(data: (C, T) array, params: argparse.Namespace) -> (C', T') array

(ref: CONFIG.md:25-34, call site preprocess/preprocessor.py:58-59), executed on the B200.

``data`` may be a numpy array (copied to the device, result copied back with the
reference's dtype conventions) or a CUDA tensor (result stays on
the device as float32 -- this is what the step runner uses to keep a recording resident
between steps).  Parameter names, defaults, side effects on ``params`` and error types
mirror the reference modules cited on each function.
"""
from __future__ import annotations

from argparse import Namespace
from typing import Callable, Dict

import numpy as np
import torch

from . import ops
from . import runtime as rt


def _wrap(fn: Callable, data, params, ref_dtype):
    """Run a device op for either kind of caller."""
    if rt.is_device(data):
        return fn(ops.as_signal(data))
    data = np.asarray(data)
    if data.ndim != 2:
        raise ValueError(f"expected a (channels, time) array, got shape {data.shape}")
    in_dtype = data.dtype
    y = fn(rt.to_device(data))
    return rt.to_host(y, rt.output_dtype(ref_dtype(in_dtype)))


_same = lambda dt: dt if np.issubdtype(dt, np.floating) else np.float64
_f64 = lambda dt: np.float64


# ------------------------------------------------------------ frequency_filter
def band_plan(params: Namespace):
    """Validated ``[(method, params dict), ...]`` of a frequency_filter step (ref: frequency_filter.py:33-74:
    same checks, same messages; unknown methods are silently ignored like the reference does)."""
    bands = getattr(params, "bands", None)
    if bands is None:
        raise ValueError("bands must be specified in params.")
    plan = []
    for cfg in bands:
        method = cfg.get("method", "hilbert")
        p = dict(cfg.get("params", {}) or {})
        if method == "hilbert":
            if "freq_ranges" not in p:
                raise ValueError("Hilbert filter requires 'freq_ranges' in params.")
        elif method == "butter":
            if "freqs" not in p:
                raise ValueError("Butterworth filter requires 'freq_range' in params.")
        elif method == "fir":
            if "order" not in p or "center_frequencies" not in p:
                raise ValueError("FIR filter requires 'order' and 'center_frequencies' in params.")
        else:
            continue        # the reference silently ignores unknown methods (:42-74)
        plan.append((method, p))
    if not plan:
        raise ValueError("need at least one array to concatenate")
    return plan


def frequency_filter(data, params: Namespace):
    """ref: preprocess/signal/frequency_filter.py:9-77.  ``params.bands`` is a list of
    {method: hilbert|butter|fir, params: {...}}; every band filters the SAME input and the
    results are concatenated along the channel axis."""
    plan = band_plan(params)
    fs = params.signal_freq

    def run_band(x, method, p, out=None):
        if method == "hilbert":
            return ops.hilbert(x, fs, out=out, **p)
        if method == "butter":
            return ops.butter(x, fs=fs, out=out, **p)
        return ops.fir_bank(x, fs, p["order"], p["center_frequencies"], out=out)

    def fn(x):
        if len(plan) == 1:
            return run_band(x, *plan[0])
        Cn, T = x.shape
        y = torch.empty((Cn * len(plan), T), dtype=torch.float32, device=x.device)
        for i, (method, p) in enumerate(plan):
            run_band(x, method, p, out=y[i * Cn:(i + 1) * Cn])
        return y

    # hilbert / butter promote to float64 in the reference; an all-fir list keeps the input dtype
    ref = _f64 if any(m != "fir" for m, _ in plan) else _same
    return _wrap(fn, data, params, ref)


# ---------------------------------------------------------------- car_rereference
def car_rereference(data, params: Namespace):
    """ref: preprocess/signal/car_rereference.py:5-41 (sets params.exclude_channels=[] when absent)."""
    excl = car_exclusions(params, data.shape[0])
    return _wrap(lambda x: ops.car(x, excl), data, params, _same)


def car_exclusions(params: Namespace, n_ch: int) -> list:
    """Validated ``exclude_channels`` (ref: car_rereference.py:23-32, including the side effect on params)."""
    if not hasattr(params, "exclude_channels"):
        params.exclude_channels = []
    excl = params.exclude_channels
    if not isinstance(excl, list):
        raise ValueError("exclude_channels must be a list of integers.")
    if any(ch < 0 or ch >= n_ch for ch in excl):
        raise ValueError("exclude_channels contains invalid channel indices.")
    return excl


# ----------------------------------------------------------------- channel_zscore
def channel_zscore(data, params: Namespace):
    """ref: preprocess/signal/channel_zscore.py:5-29 (population std; NaN kept unless preserve_nans=False)."""
    keep = getattr(params, "preserve_nans", True)
    return _wrap(lambda x: ops.zscore(x, nan_to_zero=not keep), data, params, _same)


# ------------------------------------------------------------- zscore_rereference
def zscore_rereference(data, params: Namespace):
    """ref: preprocess/signal/zscore_rereference.py:6-30,33-70."""
    if not hasattr(params, "rereference_interval") or not hasattr(params, "signal_freq"):
        raise ValueError("params must have 'rereference_interval' and 'signal_freq' attributes.")
    try:
        start, end = params.rereference_interval
    except (ValueError, TypeError):
        raise ValueError("reference_time must be a tuple of (start, end)")
    s = int(start * params.signal_freq)          # float64 product, truncation (:25-26)
    e = int(end * params.signal_freq)
    if s < 0 or e > data.shape[1]:
        raise ValueError("Reference time indices are out of bounds.")
    if s >= e:
        raise ValueError("Start time must be less than end time.")
    return _wrap(lambda x: ops.zscore(x, s, e), data, params, _same)


# ----------------------------------------------------------------- rolling_zscore
def rolling_zscore(data, params: Namespace):
    """ref: preprocess/signal/rolling_zscore.py:5-49 (trailing window, min_periods=1, ddof=1)."""
    window_length = getattr(params, "window_length", 10)
    window = int(window_length * params.signal_freq)
    keep = getattr(params, "preserve_nans", True)
    if window <= 1:
        raise ValueError("window_size must be greater than 1.")
    return _wrap(lambda x: ops.rolling_zscore(x, window, nan_to_zero=not keep), data, params, _f64)


# --------------------------------------------------------------------- downsample
def downsample(data, params: Namespace):
    """ref: preprocess/signal/downsample.py:6-29.  ``num = int(T * (target / fs))`` in float64;
    ``params.signal_freq`` becomes the requested target (CONFIG.md:34)."""
    target = getattr(params, "downsample_freq", 400)
    factor = target / params.signal_freq
    num = int(data.shape[1] * factor)
    out = _wrap(lambda x: ops.fft_resample(x, num), data, params, _same)
    params.signal_freq = target
    return out


STEPS: Dict[str, Callable] = {
    "frequency_filter": frequency_filter,
    "car_rereference": car_rereference,
    "channel_zscore": channel_zscore,
    "zscore_rereference": zscore_rereference,
    "rolling_zscore": rolling_zscore,
    "downsample": downsample,
}

# dtype a numpy caller of the reference would hold after each step (Appendix A7)
PROMOTES_TO_F64 = {"rolling_zscore"}


def reference_dtype_after(name: str, step_params: dict, dtype_in) -> np.dtype:
    dtype_in = np.dtype(dtype_in)
    if not np.issubdtype(dtype_in, np.floating):
        dtype_in = np.dtype(np.float64)
    if name in PROMOTES_TO_F64:
        return np.dtype(np.float64)
    if name == "frequency_filter":
        methods = [b.get("method", "hilbert") for b in (step_params.get("bands") or [])]
        return np.dtype(np.float64) if any(m != "fir" for m in methods) else dtype_in
    return dtype_in
