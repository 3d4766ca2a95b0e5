"""B200-native ECoG preprocessing, epoching and channel selection.

Host code is thin Python; the arithmetic runs in ``libecog_sm100.so`` (hand-written
sm_100a CUDA kernels behind a C ABI, see ``include/ecog_sm100.h``).  There is no CPU
fallback: importing an op without the built library raises.
"""
__version__ = "0.1.0"

import os as _os
import sys as _sys

DROPIN_DIR = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "dropin")


def install_dropin() -> str:
    """Put the drop-in tree first on ``sys.path`` so that the reference's dotted names
    (``preprocess.downsample``, ``preprocess.signal.downsample``, ``extract_samples``,
    ``channel_selection.active``, ``main`` ...) resolve to this package."""
    if DROPIN_DIR in _sys.path:
        _sys.path.remove(DROPIN_DIR)
    _sys.path.insert(0, DROPIN_DIR)
    # forget same-named modules that were imported from somewhere else (e.g. a reference checkout)
    roots = {"preprocess", "utils", "data_loading", "channel_selection", "main", "preprocess_main",
             "extract_samples", "channel_selection_main"}
    for name in list(_sys.modules):
        if name.split(".")[0] in roots:
            mod = _sys.modules[name]
            origin = getattr(mod, "__file__", None) or next(iter(getattr(mod, "__path__", []) or []), "")
            if not str(origin).startswith(DROPIN_DIR):
                del _sys.modules[name]
    return DROPIN_DIR
