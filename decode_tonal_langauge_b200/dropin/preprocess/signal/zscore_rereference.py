"""Drop-in module: same dotted name and entry points as the reference's `preprocess/signal/zscore_rereference.py`;
the implementation lives in decode_tonal_langauge_b200 and runs on the B200."""
from decode_tonal_langauge_b200.steps import zscore_rereference as run  # noqa: F401


def rereference(data, reference_time):
    """ref: zscore_rereference.py:33-70 (indices, not seconds)."""
    from argparse import Namespace
    try:
        start, end = reference_time
    except (TypeError, ValueError):
        raise ValueError("reference_time must be a tuple of (start, end)")
    return run(data, Namespace(signal_freq=1, rereference_interval=[start, end]))
