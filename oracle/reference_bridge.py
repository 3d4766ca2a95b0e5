"""Import the UNMODIFIED reference from /root/reference (build container only).

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  /root/reference does not exist
on the GPU box, so nothing in the ``-m gpu`` tests, ``smoke()`` or ``bench.py``
may call this; it is used by ``tests/golden/make_golden.py`` (which wrote the
committed fixtures) and by ``tests/test_oracle_vs_reference.py`` (skipped when
the tree is absent).  Four absent third-party modules are stubbed in
``sys.modules``; no reference source is copied.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("ECOG_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "preprocess", "signal"))


def _stub(name: str, **attrs):
    if name not in sys.modules:
        try:
            importlib.import_module(name)
            return
        except Exception:
            pass
        mod = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(mod, k, v)
        sys.modules[name] = mod


def load(module: str):
    """Return a reference module, e.g. ``load('preprocess.signal.downsample')``."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _stub("matplotlib")
    _stub("matplotlib.pyplot")
    _stub("textgrid", TextGrid=type("TextGrid", (), {}))
    _stub("tdt")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # the product ships drop-in modules with the same dotted names; make sure the
    # reference's own files win inside this helper
    for name in list(sys.modules):
        root = name.split(".")[0]
        if root in ("preprocess", "data_loading", "channel_selection", "utils",
                    "extract_samples", "main", "preprocess_main",
                    "channel_selection_main"):
            f = getattr(sys.modules[name], "__file__", None) or ""
            paths = list(getattr(sys.modules[name], "__path__", []) or [])
            if f and not f.startswith(REFERENCE_ROOT):
                del sys.modules[name]
            elif not f and paths and not any(p.startswith(REFERENCE_ROOT) for p in paths):
                del sys.modules[name]
    sys.path.remove(REFERENCE_ROOT)
    sys.path.insert(0, REFERENCE_ROOT)
    return importlib.import_module(module)
