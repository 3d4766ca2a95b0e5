"""Drop-in module: same dotted name and entry points as the reference's `preprocess/signal/frequency_filter.py`;
the implementation lives in decode_tonal_langauge_b200 and runs on the B200."""
from decode_tonal_langauge_b200.steps import frequency_filter as run  # noqa: F401
from decode_tonal_langauge_b200 import ops as _ops, runtime as _rt
import numpy as _np


def _host(fn, data, ref_dtype):
    if _rt.is_device(data):
        return fn(data)
    return _rt.to_host(fn(_rt.to_device(_np.asarray(data))), _rt.output_dtype(ref_dtype))


def hilbert_filter(data, sampling_rate, freq_ranges, **kw):
    return _host(lambda x: _ops.hilbert(x, sampling_rate, freq_ranges, **kw), data, _np.float64)


def butter_filter(data, freqs, fs, order=4, causal=False, filter_type="bandpass"):
    return _host(lambda x: _ops.butter(x, freqs, fs, order, causal, filter_type), data, _np.float64)
