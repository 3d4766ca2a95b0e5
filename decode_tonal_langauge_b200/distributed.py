"""Multi-GPU plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink).

The path shards two ways:

* **sessions / blocks** are independent (ref: preprocess/pipelines/subject_block.py:74-101 loops
  blocks with no cross-block state): ``assign_sessions`` deals them round-robin, no collective;
* **channels** of one recording: every step is row-independent except CAR, whose only
  cross-channel quantity is the per-timestep sum over the included channels
  (ref: preprocess/signal/car_rereference.py:34-39).  Each rank owns a contiguous block of
  rows, ``ecog_car_colsum`` produces its partial sums, ONE ``all_reduce(SUM)`` of T floats
  runs between the two kernel enqueues, and ``ecog_car_apply`` subtracts the global mean.
  Channel scores / selected sets are all-gathered at the end.

``backend`` is the object providing ``car_colsum`` / ``car_apply`` (default: the CUDA ops).
The CPU test-suite injects a numpy backend to exercise this file over ``gloo``; the product
never does.
"""
from __future__ import annotations

import os
from argparse import Namespace
from copy import deepcopy
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


OVERLAP_GROUPS = 8          # time-tile groups of the overlapped CAR all-reduce (5.4 MB of column sums each at C4)
_comm_streams = {}


def _comm_stream(device):
    key = str(device)
    if key not in _comm_streams:
        _comm_streams[key] = torch.cuda.Stream(device=device)
    return _comm_streams[key]


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of ``range(n)``: the first ``n % world`` ranks get one extra."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def assign_sessions(n_sessions: int, rank: int, world: int) -> List[int]:
    """Round-robin deal of independent sessions (no communication)."""
    return list(range(rank, n_sessions, world))


def _default_backend():
    from . import ops
    return ops


def local_exclusions(exclude_channels: Sequence[int], c_lo: int, rows_local: int, n_channels: int, bands: int = 1):
    """Validated global ``exclude_channels`` -> (local row indices, global included count).

    After a multi-band ``frequency_filter`` the reference's array is band-major: global row
    ``b * n_channels + c`` (ref: frequency_filter.py:75 concatenates the bands along the channel axis) and
    CAR spans all ``bands * n_channels`` rows.  A shard holds rows ``c_lo .. c_lo + rows_local / bands`` of
    every band, band-major as well, so global row (b, c) is local row ``b * (rows_local / bands) + c - c_lo``."""
    if not isinstance(exclude_channels, (list, tuple)):
        raise ValueError("exclude_channels must be a list of integers.")
    total = bands * n_channels
    if any(ch < 0 or ch >= total for ch in exclude_channels):
        raise ValueError("exclude_channels contains invalid channel indices.")
    per = rows_local // bands
    local = []
    for ch in exclude_channels:
        b, c = divmod(int(ch), n_channels)
        if c_lo <= c < c_lo + per:
            local.append(b * per + c - c_lo)
    return local, total - len(set(exclude_channels))


def car_sharded(x_local: torch.Tensor, c_lo: int, n_channels: int, exclude_channels: Sequence[int] = (),
                group=None, backend=None, reduce=None, bands: int = 1) -> torch.Tensor:
    """CAR of a channel shard.  ``x_local`` holds global channels [c_lo, c_lo + rows / bands) of each of the
    ``bands`` concatenated band copies (``bands`` = 1: plain rows [c_lo, c_lo + rows))."""
    backend = backend or _default_backend()
    local_excl, n_included = local_exclusions(exclude_channels, c_lo, x_local.shape[0], n_channels, bands)
    weights = None
    if local_excl:
        w = np.ones(x_local.shape[0], dtype=np.float32)
        w[local_excl] = 0.0
        weights = torch.from_numpy(w).to(x_local.device)
    partial = backend.car_colsum(x_local, weights)
    if reduce is not None:
        reduce(partial)
    elif dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM, group=group)     # T floats over NVLink
    return backend.car_apply(x_local, partial, n_included)


def preprocess_signal_sharded(x_local, steps: List[dict], block_params: Namespace, c_lo: int, n_channels: int,
                              group=None, backend=None, fuse: Optional[bool] = None, timing: Optional[list] = None,
                              profile: Optional[list] = None, overlap_allreduce: Optional[bool] = None):
    """``preprocess_signal`` for a channel shard resident on this rank's device: identical step
    semantics (same shared parameter Namespace, same fusion groups), except that the CAR column
    sums are all-reduced between the two kernel enqueues.  Returns ``(local tensor, signal_freq,
    bands)`` where ``bands`` is the number of concatenated band copies the local rows are organised
    in (see ``gather_channels``).  ``timing``: optional list; a (start, end) CUDA-event pair is appended
    around every all-reduce (recorded on the current stream); ``profile``: optional list, one (group name,
    start, end) CUDA-event triple per executed group (the all-reduce lies inside its CAR group)."""
    from . import preprocessor as P
    x = x_local
    bands = 1
    if overlap_allreduce is None:
        # measured (profiles/r02_bench_n{2,4,8}.json, C4): the 43 MB collective takes 0.2-0.4 ms as one call;
        # split into 8 time-tile groups on a side stream it is hidden, but the 8 partial launches and stream
        # hops cost more than that at N >= 4 (67.2 vs 65.2 ms at N = 4, 35.2 vs 34.3 ms at N = 8; -0.1 ms at N = 2)
        overlap_allreduce = os.environ.get("ECOG_OVERLAP_ALLREDUCE", "0") == "1"
    plain = backend is not None            # the CPU test-suite's numpy backend has no fused kernels
    groups = P.fusion_groups(steps) if (P.fusion_enabled(fuse) and not plain) else [("step", s) for s in steps]

    def reduce(colsum):
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
            return
        if timing is not None and colsum.is_cuda:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dist.all_reduce(colsum, op=dist.ReduceOp.SUM, group=group)
            e1.record()
            timing.append((e0, e1))
        else:
            dist.all_reduce(colsum, op=dist.ReduceOp.SUM, group=group)

    def overlap(xl, w, n_inc, fs, hp):
        """CAR column sums -> all-reduce -> Hilbert by TIME-TILE GROUPS: the column sums of group g are
        all-reduced on a side stream while the kernels of the other groups run; a group's Hilbert blocks
        start as soon as the sums of its own samples and of the two neighbouring groups (the circular
        halo) have arrived.  Declines (None) on one rank, on CPU tensors and for whole-record banks."""
        from . import ops
        if not (xl.is_cuda and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
            return None
        info = ops.hilbert_block_info(xl.shape[1], fs, **hp)
        if info is None:
            return None
        halo, U, nblk = info
        G = min(OVERLAP_GROUPS, nblk // 4)
        if G < 3:
            return None
        T = xl.shape[1]
        cuts = [2 * ((nblk * g // G) // 2) for g in range(G)] + [nblk]          # even block boundaries
        main = torch.cuda.current_stream()
        comm = _comm_stream(xl.device)
        colsum = torch.empty(T, dtype=torch.float32, device=xl.device)
        colsum.record_stream(comm)
        done = []
        for g in range(G):
            s0, s1 = cuts[g] * U, min(cuts[g + 1] * U, T)
            ops.car_colsum(xl[:, s0:s1], w, out=colsum[s0:s1])
            ready = torch.cuda.Event()
            ready.record(main)
            with torch.cuda.stream(comm):
                comm.wait_event(ready)
                e0, e1 = torch.cuda.Event(enable_timing=timing is not None), torch.cuda.Event(enable_timing=timing is not None)
                e0.record(comm)
                dist.all_reduce(colsum[s0:s1], op=dist.ReduceOp.SUM, group=group)
                e1.record(comm)
            done.append(e1)
            if timing is not None:
                timing.append((e0, e1))
        y = torch.empty((xl.shape[0], T), dtype=torch.float32, device=xl.device)
        waited = set()
        for g in list(range(1, G)) + [0]:             # group 0 needs the LAST sums (circular halo): it goes last
            for j in ((g - 1) % G, g, (g + 1) % G):
                if j not in waited:
                    main.wait_event(done[j])
                    waited.add(j)
            ops.hilbert(xl, fs, car=(colsum, n_inc), out=y, blocks=(cuts[g], cuts[g + 1]), **hp)
        return y

    for g in groups:
        if profile is not None and x.is_cuda:
            ev0 = torch.cuda.Event(enable_timing=True)
            ev0.record()
        if g[0] == "step" and P._short(g[1]) == "car_rereference":
            P.apply_step_params(block_params, g[1])
            excl = getattr(block_params, "exclude_channels", None)
            if excl is None:
                excl = block_params.exclude_channels = []
            # after a multi-band frequency_filter the global layout is band-major; CAR then spans
            # bands * n_channels rows and this shard holds `bands` slices of them (local_exclusions)
            x = car_sharded(x, c_lo, n_channels, excl, group, backend, reduce=reduce, bands=bands)
        elif g[0] == "car_hilbert":
            x = P._run_group(x, g, block_params, False,
                             shard=(c_lo, n_channels, reduce, bands, overlap if overlap_allreduce else None))
        else:
            x = P._run_group(x, g, block_params, False)
            if g[0] == "step" and P._short(g[1]) == "frequency_filter":
                bands *= max(1, len(getattr(block_params, "bands", []) or []))
        if profile is not None and x.is_cuda:
            ev1 = torch.cuda.Event(enable_timing=True)
            ev1.record()
            profile.append((P._group_name(g), ev0, ev1))
    return x, block_params.signal_freq, bands


def gather_channels(y_local: torch.Tensor, n_channels: int, bands: int = 1, group=None) -> torch.Tensor:
    """All-gather shard outputs into the reference's global row order.  A multi-band
    ``frequency_filter`` concatenates band-major, so rank r's rows [b*Cr, (b+1)*Cr) go to global
    rows b*C + [c_lo, c_hi)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return y_local
    T = y_local.shape[1]
    sizes = [shard_bounds(n_channels, r, world) for r in range(world)]
    cmax = max(hi - lo for lo, hi in sizes) * bands
    pad = torch.zeros((cmax, T), dtype=y_local.dtype, device=y_local.device)
    pad[: y_local.shape[0]] = y_local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    out = torch.empty((bands * n_channels, T), dtype=y_local.dtype, device=y_local.device)
    for r, (lo, hi) in enumerate(sizes):
        cr = hi - lo
        for b in range(bands):
            out[b * n_channels + lo: b * n_channels + hi] = parts[r][b * cr:(b + 1) * cr]
    return out


def gather_selection(runs_local: torch.Tensor, c_lo: int, n_channels: int, length_threshold: int,
                     group=None) -> List[int]:
    """All-gather per-shard longest-run vectors ((C/P,) int32) and apply the strict
    ``run > length_threshold`` rule (ref: channel_selection/utils.py:73) on the global vector."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        full = runs_local
    else:
        sizes = [shard_bounds(n_channels, r, world) for r in range(world)]
        cmax = max(hi - lo for lo, hi in sizes)
        pad = torch.zeros(cmax, dtype=runs_local.dtype, device=runs_local.device)
        pad[: runs_local.shape[0]] = runs_local
        parts = [torch.empty_like(pad) for _ in range(world)]
        dist.all_gather(parts, pad, group=group)
        full = torch.cat([parts[r][: hi - lo] for r, (lo, hi) in enumerate(sizes)])
    return [int(c) for c in torch.nonzero(full > length_threshold).flatten().cpu().tolist()]
