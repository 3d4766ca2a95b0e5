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


# ------------------------------------------------------------------ pipelined host path
PIPELINE_MIN_CHANNELS = 64       # below this the recording is moved and processed in one piece
PIPELINE_CHUNKS = int(os.environ.get("ECOG_PIPELINE_CHUNKS", "12"))   # measured at C2: 8 -> 195.4 ms, 12 -> 193.9, 16 -> 193.9


def _row_independent(name: str, step_params: dict) -> bool:
    """Steps whose output row c depends on input row c only and that keep the row count."""
    if name == "frequency_filter":
        return len(step_params.get("bands") or []) == 1
    return name in ("channel_zscore", "zscore_rereference", "rolling_zscore", "downsample")


def _scoped(step: dict, block_params: Namespace) -> Namespace:
    scope = Namespace(**vars(block_params))
    for key, value in (step.get("params", {}) or {}).items():
        setattr(scope, key, deepcopy(value))
    return scope


def _run_segment(x, segment, block_params: Namespace):
    """Run consecutive steps on a device tensor with a private copy of the parameters; returns
    (tensor, params after the segment)."""
    params = Namespace(**vars(block_params))
    for step in segment:
        scope = _scoped(step, params)
        _, fn = _resolve(step["module"])
        x = fn(x, scope)
        params.signal_freq = scope.signal_freq
    return x, params


def _pipelined_host_run(data: np.ndarray, steps: List[dict], block_params: Namespace, out_dtype):
    """Host array in, host array out, with the PCIe copies hidden behind the kernels.

    Every step but ``car_rereference`` is row independent (ref: each preprocess/signal/*.py works
    along axis 1), so the recording is cut into channel chunks: chunk i+1 is copied to the device
    while the steps BEFORE the first CAR run on chunk i, and the steps AFTER the last CAR run on
    chunk i while chunk i-1 travels back.  CAR itself (and anything between two CARs) sees the
    whole array.  Returns None when the step list does not have that shape."""
    import torch
    names = [s["module"].split(".")[-1] for s in steps]
    if data.ndim != 2 or data.shape[0] < PIPELINE_MIN_CHANNELS or not all(n in S.STEPS for n in names):
        return None
    ok = [n == "car_rereference" or _row_independent(n, s.get("params", {}) or {}) for n, s in zip(names, steps)]
    if not all(ok):
        return None
    cars = [i for i, n in enumerate(names) if n == "car_rereference"]
    head = steps[:cars[0]] if cars else []
    middle = steps[cars[0]:cars[-1] + 1] if cars else []
    tail = steps[cars[-1] + 1:] if cars else steps
    Cn, T = data.shape
    bounds = np.linspace(0, Cn, PIPELINE_CHUNKS + 1).astype(int)
    chunks = [(int(a), int(b)) for a, b in zip(bounds[:-1], bounds[1:]) if b > a]
    dev = rt.device()
    main = torch.cuda.current_stream()
    up, down = _copy_streams(dev)
    up.wait_stream(main)
    down.wait_stream(main)
    src = torch.from_numpy(data)
    pinned = src.is_pinned()
    x0 = torch.empty((Cn, T), dtype=src.dtype, device=dev)
    arrived = []
    with torch.cuda.stream(up):
        for a, b in chunks:
            x0[a:b].copy_(src[a:b], non_blocking=pinned)
            ev = torch.cuda.Event()
            ev.record(up)
            arrived.append(ev)
    rt.h2d_bytes += src.numel() * src.element_size()
    # head: per chunk, as soon as it has arrived
    params = block_params
    waited = [False] * len(chunks)

    def chunk_in(i):
        if not waited[i]:
            main.wait_event(arrived[i])
            waited[i] = True
        a, b = chunks[i]
        xc = x0[a:b]
        return xc if xc.dtype == torch.float32 else xc.to(torch.float32)

    if cars:
        x1 = None
        for i, (a, b) in enumerate(chunks):
            yc = chunk_in(i)
            if head:
                yc, params = _run_segment(yc, head, block_params)
            if x1 is None:
                x1 = x0 if (not head and x0.dtype == torch.float32) else \
                    torch.empty((Cn, yc.shape[1]), dtype=torch.float32, device=dev)
            if x1 is not x0:
                x1[a:b].copy_(yc)
        x1, params = _run_segment(x1, middle, params)
        chunk_mid = lambda i: x1[chunks[i][0]:chunks[i][1]]
    else:
        chunk_mid = chunk_in              # no CAR: the whole list runs per chunk, both copies overlap it
    # tail: per chunk, result cast on the device and sent back while the next chunk is computed
    td = getattr(torch, np.dtype(out_dtype).name)
    out = None
    keep = []
    final = params
    for i, (a, b) in enumerate(chunks):
        yc = chunk_mid(i)
        if tail:
            yc, final = _run_segment(yc, tail, params)
        yc = yc.to(td) if yc.dtype != td else yc
        if out is None:
            out = torch.empty((Cn, yc.shape[1]), dtype=td, pin_memory=True)
        ev = torch.cuda.Event()
        ev.record(main)
        down.wait_event(ev)
        with torch.cuda.stream(down):
            out[a:b].copy_(yc, non_blocking=True)
        keep.append(yc)
    down.synchronize()
    main.synchronize()
    rt.d2h_bytes += out.numel() * out.element_size()
    block_params.signal_freq = final.signal_freq
    return out.numpy()


_streams = {}


def _copy_streams(dev):
    import torch
    key = str(dev)
    if key not in _streams:
        _streams[key] = (torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev))
    return _streams[key]


def preprocess_signal(data, steps: List[dict], block_params: Namespace, figure_dir: Optional[str] = None,
                      num_channels: int = 5, duration: float = 1.0, strict_params: bool = False,
                      profile: Optional[list] = None):
    """``profile``: optional list; one (step name, start event, end event) CUDA-event triple
    per step is appended (events recorded on the current stream, no synchronisation)."""
    was_host = not rt.is_device(data)
    ref_dtype = np.asarray(data).dtype if was_host else np.dtype(np.float32)
    if was_host and len(steps) and not strict_params and profile is None and os.environ.get("ECOG_PIPELINE", "1") != "0":
        dt = ref_dtype
        for step in steps:
            dt = S.reference_dtype_after(step["module"].split(".")[-1], step.get("params", {}) or {}, dt)
        arr = np.asarray(data)
        if np.issubdtype(arr.dtype, np.floating) and arr.flags.c_contiguous:
            if not arr.flags.writeable:
                arr = arr.copy()
            y = _pipelined_host_run(arr, steps, block_params, rt.output_dtype(dt))
            if y is not None:
                return y, block_params.signal_freq
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
