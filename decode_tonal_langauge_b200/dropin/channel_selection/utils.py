"""Drop-in module: same dotted name and entry points as the reference's `channel_selection/utils.py`;
the implementation lives in decode_tonal_langauge_b200 and runs on the B200."""
import numpy as _np


def get_max_length(indices) -> int:
    """Longest stretch of consecutive integers in a sorted index array (ref: channel_selection/utils.py:4-30)."""
    idx = _np.asarray(indices)
    if idx.size == 0:
        raise IndexError("index 0 is out of bounds for axis 0 with size 0")
    breaks = _np.flatnonzero(_np.diff(idx) != 1)
    edges = _np.concatenate(([-1], breaks, [idx.size - 1]))
    return int(_np.diff(edges).max())


def find_significant_channels(p_values, pvalue_threshold: float = 0.05, length_threshold: int = 10):
    """ref: channel_selection/utils.py:33-76; the run-length scan runs on the device."""
    from decode_tonal_langauge_b200 import ops, runtime
    p = runtime.to_device(_np.asarray(p_values, dtype=_np.float64), dtype=None)
    runs = ops.sig_runlength(p, pvalue_threshold / p.shape[1]).cpu().numpy()
    return [int(c) for c in _np.nonzero(runs > length_threshold)[0]], []
