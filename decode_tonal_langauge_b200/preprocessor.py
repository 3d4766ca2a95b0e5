"""Operator host: the reference's ``preprocess/preprocessor.py`` contract on the B200.

``preprocess_signal(data, steps, block_params, figure_dir=None)`` runs a YAML ``steps``
list (ref: preprocess/preprocessor.py:39-70) and returns ``(data, signal_freq)``.
Differences, all supersets:
  * the recording is copied to the device ONCE, stays resident between steps and is
    copied back ONCE in the dtype the reference would have produced;
  * every step gets a fresh parameter scope (the reference's shared Namespace forbids
    repeating a step, B6); ``strict_params=True`` restores the reference's collision error;
  * step modules are resolved by their last dotted component, so ``preprocess.<name>``
    (docs, example_config.yaml) and ``preprocess.signal.<name>`` (real location) both work
    (B1); unknown modules are imported and called like the reference does.
"""
from __future__ import annotations

import importlib
import os
from argparse import Namespace
from copy import deepcopy
from typing import List, Optional

import numpy as np

from . import runtime as rt
from . import steps as S


def _resolve(module_name: str):
    short = module_name.split(".")[-1]
    if short in S.STEPS:
        return short, S.STEPS[short]
    module = importlib.import_module(module_name)      # user-supplied step (ref :58)
    return short, module.run


def _profile_name(name: str, step_params: dict) -> str:
    if name == "frequency_filter":
        kinds = []
        for b in step_params.get("bands") or []:
            m = b.get("method", "hilbert")
            kinds.append(m if m != "butter" else "butter_" + str((b.get("params") or {}).get("filter_type", "bandpass")))
        return "frequency_filter[" + "+".join(kinds) + "]"
    return name


def preprocess_signal(data, steps: List[dict], block_params: Namespace, figure_dir: Optional[str] = None,
                      num_channels: int = 5, duration: float = 1.0, strict_params: bool = False,
                      profile: Optional[list] = None):
    """``profile``: optional list; one (step name, start event, end event) CUDA-event triple
    per step is appended (events recorded on the current stream, no synchronisation)."""
    was_host = not rt.is_device(data)
    ref_dtype = np.asarray(data).dtype if was_host else np.dtype(np.float32)
    if was_host and len(steps) and all(s["module"].split(".")[-1] in S.STEPS for s in steps):
        x = rt.to_device(np.asarray(data))
    else:
        x = data
    for step in steps:
        step_params = step.get("params", {}) or {}
        if strict_params:
            for key in step_params:
                if hasattr(block_params, key):
                    raise ValueError(
                        f"Parameter '{key}' already exists in params. "
                        "Please ensure no conflicting parameter names in each preprocessing step.")
        scope = block_params if strict_params else Namespace(**vars(block_params))
        for key, value in step_params.items():
            setattr(scope, key, deepcopy(value))
        name, fn = _resolve(step["module"])
        if profile is not None and rt.is_device(x):
            import torch
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            x = fn(x, scope)
            ev1.record()
            profile.append((_profile_name(name, step_params), ev0, ev1))
        else:
            x = fn(x, scope)
        block_params.signal_freq = scope.signal_freq
        ref_dtype = S.reference_dtype_after(name, step_params, ref_dtype)
    if was_host and rt.is_device(x):
        x = rt.to_host(x, rt.output_dtype(ref_dtype))
    return x, block_params.signal_freq


def preprocess_modalities(data_dict: dict, modalities_cfg: dict, base_params: Namespace,
                          figure_dir: Optional[str] = None) -> dict:
    """ref: preprocess/preprocessor.py:8-36 (figure output is not produced; B7 guarded)."""
    for modality, cfg in modalities_cfg.items():
        if cfg.get("type") is None:
            raise KeyError(f"Modality '{modality}' missing 'type' field in config")
        if figure_dir:
            os.makedirs(os.path.join(figure_dir, modality), exist_ok=True)
        steps = (cfg.get("preprocessing") or {}).get("steps", [])
        if not steps:
            continue
        if cfg.get("type") != "signal":
            continue
        params = deepcopy(base_params)
        params.signal_freq = data_dict.get(f"{modality}_sf")
        processed, freq = preprocess_signal(data_dict[modality], steps, params)
        if freq is not None:
            data_dict[f"{modality}_sf"] = freq
        data_dict[modality] = processed
    return data_dict
