"""Run a YAML-style ``steps`` list through the oracle operators.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Mirrors the dispatch loop of
ref: preprocess/preprocessor.py:39-70 with a fresh parameter scope per step
 and ``signal_freq`` threaded through.
"""
from __future__ import annotations

import numpy as np

from . import steps as S


def run_step(name: str, data, fs, params: dict):
    name = name.split(".")[-1]
    if name == "frequency_filter":
        return S.frequency_filter(data, fs, params.get("bands")), fs
    if name == "car_rereference":
        return S.car_rereference(data, params.get("exclude_channels", [])), fs
    if name == "channel_zscore":
        return S.channel_zscore(data, params.get("preserve_nans", True)), fs
    if name == "zscore_rereference":
        return S.zscore_rereference(data, fs, params["rereference_interval"]), fs
    if name == "rolling_zscore":
        return S.rolling_zscore(data, fs, params.get("window_length", 10),
                                params.get("preserve_nans", True)), fs
    if name == "downsample":
        return S.downsample(data, fs, params.get("downsample_freq", 400))
    raise KeyError(name)


def run_chain(data: np.ndarray, fs, steps):
    for step in steps:
        data, fs = run_step(step["module"], data, fs, step.get("params", {}) or {})
    return data, fs
