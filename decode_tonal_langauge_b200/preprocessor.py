"""Operator host: the reference's ``preprocess/preprocessor.py`` contract on the B200.

``preprocess_signal(data, steps, block_params, figure_dir=None)`` runs a YAML ``steps``
list (ref: preprocess/preprocessor.py:39-70) and returns ``(data, signal_freq)``.
Differences, all supersets:
  * the recording is copied to the device ONCE, stays resident between steps and is
    copied back ONCE in the dtype the reference would have produced (``output_dtype``
    overrides: float32 halves the return traffic, widening is then the caller's choice);
  * step parameters are written into ONE shared Namespace exactly like the reference does
    (ref :46-53), so a key set by an earlier step stays visible to the later ones
    (``preserve_nans``, ``exclude_channels`` ...); a repeated key is overwritten (last write
    wins) instead of raising, because the reference's collision error forbids repeating a step
    (B6); ``strict_params=True`` restores that error;
  * step modules are resolved by their last dotted component, so ``preprocess.<name>``
    (docs, example_config.yaml) and ``preprocess.signal.<name>`` (real location) both work
    (B1); unknown modules are imported and called like the reference does;
  * the runner owns two re-orderings that leave the result unchanged (``fuse``):
      - ``butter(zero-phase) -> [car_rereference] -> butter(zero-phase)``: CAR is the same linear
        combination of rows at every sample and filtfilt (odd padding and ``zi * x[0]`` included)
        is linear and identical for every row, so CAR commutes with it; the two filtfilts then run
        as ONE forward and ONE backward sweep of an 8-section cascade (``ops.sosfilt_pair``:
        4 sweeps -> 2, exact edges recomputed);
      - ``car_rereference -> frequency_filter[hilbert]``: the column mean is produced by one
        read-only pass and subtracted in the Hilbert kernel's load (no read + write of the
        recording for CAR).
    The step modules keep the reference's one-step contract; ``fuse=False`` runs them one by one.
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


def _short(step: dict) -> str:
    return step["module"].split(".")[-1]


def _step_params(step: dict) -> dict:
    return step.get("params", {}) or {}


def _profile_name(name: str, step_params: dict) -> str:
    if name == "frequency_filter":
        kinds = []
        for b in step_params.get("bands") or []:
            m = b.get("method", "hilbert")
            kinds.append(m if m != "butter" else "butter_" + str((b.get("params") or {}).get("filter_type", "bandpass")))
        return "frequency_filter[" + "+".join(kinds) + "]"
    return name


def apply_step_params(params: Namespace, step: dict, strict: bool = False) -> None:
    """Flatten a step's params into the shared Namespace (ref: preprocessor.py:46-53).  The reference
    raises on a key that is already there; by default the later step wins here (B6)."""
    for key, value in _step_params(step).items():
        if strict and hasattr(params, key):
            raise ValueError(
                f"Parameter '{key}' already exists in params. "
                "Please ensure no conflicting parameter names in each preprocessing step.")
        setattr(params, key, deepcopy(value))


# --------------------------------------------------------------------------- fusion planner
def _single_band(step: dict, method: str) -> Optional[dict]:
    """The band's params dict if `step` is a frequency_filter with exactly ONE band of `method`."""
    if _short(step) != "frequency_filter":
        return None
    bands = _step_params(step).get("bands")
    if not isinstance(bands, list) or len(bands) != 1 or not isinstance(bands[0], dict):
        return None
    if bands[0].get("method", "hilbert") != method:
        return None
    return dict(bands[0].get("params", {}) or {})


def _zero_phase_butter(step: dict) -> bool:
    p = _single_band(step, "butter")
    return p is not None and "freqs" in p and not p.get("causal", False) and set(p) <= {"freqs", "order", "causal", "filter_type"}


def _plain_car(step: dict) -> bool:
    return _short(step) == "car_rereference"


def fusion_groups(steps: List[dict]) -> List[tuple]:
    """``[("step", step) | ("iir_pair", stepA, stepB) | ("car_hilbert", car_step, hilbert_step)]``.

    First CAR is moved behind a zero-phase Butterworth step that follows it when another one
    precedes it (A, CAR, B -> A, B, CAR: CAR commutes with per-row linear filters), then adjacent
    pairs are grouped."""
    order = list(steps)
    i = 0
    while i + 2 < len(order):
        if _zero_phase_butter(order[i]) and _plain_car(order[i + 1]) and _zero_phase_butter(order[i + 2]):
            order[i + 1], order[i + 2] = order[i + 2], order[i + 1]
            i += 2
        else:
            i += 1
    groups, i = [], 0
    while i < len(order):
        if i + 1 < len(order) and _zero_phase_butter(order[i]) and _zero_phase_butter(order[i + 1]):
            groups.append(("iir_pair", order[i], order[i + 1]))
            i += 2
        elif i + 1 < len(order) and _plain_car(order[i]) and _single_band(order[i + 1], "hilbert") is not None \
                and "freq_ranges" in _single_band(order[i + 1], "hilbert"):
            groups.append(("car_hilbert", order[i], order[i + 1]))
            i += 2
        else:
            groups.append(("step", order[i]))
            i += 1
    return groups


def fusion_enabled(fuse: Optional[bool]) -> bool:
    return (os.environ.get("ECOG_FUSE", "1") != "0") if fuse is None else bool(fuse)


def _group_steps(group: tuple) -> List[dict]:
    return list(group[1:])


def _group_name(group: tuple) -> str:
    names = [_profile_name(_short(s), _step_params(s)) for s in _group_steps(group)]
    return names[0] if group[0] == "step" else "+".join(names)


def _run_group(x, group: tuple, params: Namespace, strict: bool, shard=None):
    """Run one fusion group on a device tensor; ``params`` is the shared Namespace.
    ``shard`` (channel-sharded recordings, distributed.py): ``(c_lo, n_channels, reduce[, bands[, overlap]])`` --
    the local rows are global channels [c_lo, c_lo + C_local / bands) of each band copy, ``exclude_channels``
    holds global (band-major) row indices, ``reduce(colsum)`` all-reduces the column sums in place, and
    ``overlap(x, w, n_inc, fs, hilbert params)`` (optional) runs column sums, all-reduce and Hilbert blocks
    by time-tile groups with the collective on a side stream (returns None to decline)."""
    from . import design as D
    from . import ops
    kind = group[0]
    if kind == "step":
        step = group[1]
        apply_step_params(params, step, strict)
        name, fn = _resolve(step["module"])
        return fn(x, params)
    if kind == "iir_pair":
        designs = []
        for step in group[1:]:
            apply_step_params(params, step, strict)
            (_, p), = S.band_plan(params)
            designs.append(D.butter_design(p["freqs"], params.signal_freq, p.get("order", 4), False,
                                           p.get("filter_type", "bandpass")))
        return ops.sosfilt_pair(ops.as_signal(x), designs[0], designs[1])
    if kind == "car_hilbert":
        car_step, hil_step = group[1], group[2]
        apply_step_params(params, car_step, strict)
        x = ops.as_signal(x)
        if shard is None:
            excl = S.car_exclusions(params, x.shape[0])
            w, n_inc = ops._car_weights(x.shape[0], excl, x.device)
            colsum = ops.car_colsum(x, w)
        else:
            from .distributed import local_exclusions
            c_lo, n_channels, reduce = shard[:3]
            bands = shard[3] if len(shard) > 3 else 1
            excl = S.car_exclusions(params, bands * n_channels)
            local, n_inc = local_exclusions(excl, c_lo, x.shape[0], n_channels, bands)
            w, _ = ops._car_weights(x.shape[0], local, x.device)
            overlap = shard[4] if len(shard) > 4 else None
            if overlap is not None:
                apply_step_params(params, hil_step, strict)
                (_, p), = S.band_plan(params)
                y = overlap(x, w, n_inc, params.signal_freq, p)
                if y is not None:
                    return y
                colsum = ops.car_colsum(x, w)
                reduce(colsum)
                return ops.hilbert(x, params.signal_freq, car=(colsum, n_inc), **p)
            colsum = ops.car_colsum(x, w)
            reduce(colsum)
        apply_step_params(params, hil_step, strict)
        (_, p), = S.band_plan(params)
        return ops.hilbert(x, params.signal_freq, car=(colsum, n_inc), **p)
    raise ValueError(f"unknown fusion group {kind}")


# ------------------------------------------------------------------ pipelined host path
PIPELINE_MIN_CHANNELS = 64       # below this the recording is moved and processed in one piece
PIPELINE_CHUNKS = int(os.environ.get("ECOG_PIPELINE_CHUNKS", "12"))   # measured at C2: 8 -> 195.4 ms, 12 -> 193.9, 16 -> 193.9


def _row_independent(group: tuple) -> bool:
    """Groups whose output row c depends on input row c only and that keep the row count."""
    if group[0] == "iir_pair":
        return True
    if group[0] != "step":
        return False
    name, p = _short(group[1]), _step_params(group[1])
    if name == "frequency_filter":
        return len(p.get("bands") or []) == 1
    return name in ("channel_zscore", "zscore_rereference", "rolling_zscore", "downsample")


def _needs_all_rows(group: tuple) -> bool:
    return group[0] == "car_hilbert" or (group[0] == "step" and _short(group[1]) == "car_rereference")


def _run_segment(x, segment: List[tuple], block_params: Namespace):
    """Run consecutive groups on a device tensor with a private copy of the parameters; returns
    (tensor, params after the segment)."""
    params = Namespace(**vars(block_params))
    for group in segment:
        x = _run_group(x, group, params, False)
    return x, params


def _pipelined_host_run(data: np.ndarray, groups: List[tuple], block_params: Namespace, out_dtype):
    """Host array in, host array out, with the PCIe copies hidden behind the kernels.

    Every step but ``car_rereference`` is row independent (ref: each preprocess/signal/*.py works
    along axis 1), so the recording is cut into channel chunks: chunk i+1 is copied to the device
    while the groups BEFORE the first CAR run on chunk i, and the groups AFTER the last CAR run on
    chunk i while chunk i-1 travels back.  CAR itself (and anything between two CARs) sees the
    whole array; a CAR folded into the Hilbert load contributes its column-sum pass to the middle
    and its Hilbert kernel to the per-chunk tail.  Returns None when the list does not have that shape."""
    import torch
    from . import ops
    if data.ndim != 2 or data.shape[0] < PIPELINE_MIN_CHANNELS:
        return None
    for g in groups:
        if not all(_short(s) in S.STEPS for s in _group_steps(g)):
            return None
        if not (_row_independent(g) or _needs_all_rows(g)):
            return None
    cars = [i for i, g in enumerate(groups) if _needs_all_rows(g)]
    head = groups[:cars[0]] if cars else []
    middle = groups[cars[0]:cars[-1] + 1] if cars else []
    tail = groups[cars[-1] + 1:] if cars else groups
    # a trailing car_hilbert group: column sums over the whole array (middle), Hilbert per chunk (tail)
    split_last = bool(middle) and middle[-1][0] == "car_hilbert"
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
    trace = rt.trace                       # optional: where a step's time goes (bench.py, e2e.breakdown)
    import time as _time
    t_host0 = _time.perf_counter()

    def mark(label, stream):
        if trace is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record(stream)
            trace.append((label, ev, _time.perf_counter() - t_host0))

    mark("start", main)
    x0 = torch.empty((Cn, T), dtype=src.dtype, device=dev)
    arrived = []
    with torch.cuda.stream(up):
        for a, b in chunks:
            x0[a:b].copy_(src[a:b], non_blocking=pinned)
            ev = torch.cuda.Event()
            ev.record(up)
            arrived.append(ev)
    mark("h2d_done", up)
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

    car_fold = None
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
        if split_last:
            x1, params = _run_segment(x1, middle[:-1], params)
            params = Namespace(**vars(params))
            car_step, hil_step = middle[-1][1], middle[-1][2]
            apply_step_params(params, car_step)
            excl = S.car_exclusions(params, x1.shape[0])
            w, n_inc = ops._car_weights(x1.shape[0], excl, x1.device)
            colsum = ops.car_colsum(x1, w)
            apply_step_params(params, hil_step)
            (_, hp), = S.band_plan(params)
            car_fold = (colsum, n_inc, hp)
        else:
            x1, params = _run_segment(x1, middle, params)
        chunk_mid = lambda i: x1[chunks[i][0]:chunks[i][1]]
        mark("head_and_middle_done", main)
    else:
        chunk_mid = chunk_in              # no CAR: the whole list runs per chunk, both copies overlap it
    # tail: per chunk, result cast on the device and sent back while the next chunk is computed
    td = getattr(torch, np.dtype(out_dtype).name)
    out = None
    keep = []
    final = params
    for i, (a, b) in enumerate(chunks):
        yc = chunk_mid(i)
        if car_fold is not None:
            yc = ops.hilbert(yc, params.signal_freq, car=(car_fold[0], car_fold[1]), **car_fold[2])
        if tail:
            yc, final = _run_segment(yc, tail, params)
        yc = yc.to(td) if yc.dtype != td else yc
        if out is None:
            out = torch.empty((Cn, yc.shape[1]), dtype=td, pin_memory=True)
            mark("result_buffer_ready", main)
        ev = torch.cuda.Event()
        ev.record(main)
        down.wait_event(ev)
        with torch.cuda.stream(down):
            out[a:b].copy_(yc, non_blocking=True)
        keep.append(yc)
    mark("tail_done", main)
    mark("d2h_done", down)
    down.synchronize()
    main.synchronize()
    if trace is not None:
        trace.append(("host_return", None, _time.perf_counter() - t_host0))
    rt.d2h_bytes += out.numel() * out.element_size()
    for k, v in vars(final).items():          # the shared Namespace the caller handed in sees every step's keys
        setattr(block_params, k, v)
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
                      profile: Optional[list] = None, fuse: Optional[bool] = None, output_dtype=None):
    """``profile``: optional list; one (group name, start event, end event) CUDA-event triple
    per executed group is appended (events recorded on the current stream, no synchronisation).
    ``fuse``: None = on unless ECOG_FUSE=0.  ``output_dtype``: dtype of the array handed back to a
    numpy caller (None = the reference's own convention)."""
    was_host = not rt.is_device(data)
    ref_dtype = np.asarray(data).dtype if was_host else np.dtype(np.float32)
    known = all(_short(s) in S.STEPS for s in steps)
    groups = fusion_groups(steps) if (fusion_enabled(fuse) and known and not strict_params) else [("step", s) for s in steps]

    def final_dtype(dt):
        for step in steps:
            dt = S.reference_dtype_after(_short(step), _step_params(step), dt)
        return np.dtype(output_dtype) if output_dtype is not None else rt.output_dtype(dt)

    if was_host and len(steps) and not strict_params and profile is None and os.environ.get("ECOG_PIPELINE", "1") != "0":
        arr = np.asarray(data)
        if np.issubdtype(arr.dtype, np.floating) and arr.flags.c_contiguous:
            if not arr.flags.writeable:
                arr = arr.copy()
            y = _pipelined_host_run(arr, groups, block_params, final_dtype(ref_dtype))
            if y is not None:
                return y, block_params.signal_freq
    if was_host and len(steps) and known:
        x = rt.to_device(np.asarray(data))
    else:
        x = data
    for group in groups:
        if profile is not None and rt.is_device(x):
            import torch
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            x = _run_group(x, group, block_params, strict_params)
            ev1.record()
            profile.append((_group_name(group), ev0, ev1))
        else:
            x = _run_group(x, group, block_params, strict_params)
    if was_host and rt.is_device(x):
        x = rt.to_host(x, final_dtype(ref_dtype))
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
