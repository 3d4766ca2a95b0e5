"""A user-supplied CPU step module (the reference's plug-in contract) used by host tests."""
import numpy as np


def run(data, params):
    params.seen = getattr(params, "seen", 0) + 1
    gain = getattr(params, "gain", 1.0)
    if getattr(params, "halve_rate", False):
        params.signal_freq = params.signal_freq / 2
        return np.asarray(data)[:, ::2] * gain
    return np.asarray(data) * gain
