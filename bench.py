#!/usr/bin/env python
"""Benchmark of the ECoG hot path on B200 (contract: see the round brief / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C2|C1|tiny]

One "step" = one pass of the FULL6 preprocessing chain (notch -> CAR -> band-pass ->
Gaussian-Hilbert envelope -> FFT downsample -> z-score) over one synthetic session.
N = 1 runs BASELINE.json configs[1] ("C2": 256 ch x 60 min @ 2 kHz); N > 1 shards independent
sessions over the ranks (weak scaling, no data-path collective), one process per GPU under
torchrun.  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from argparse import Namespace

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (channels, samples, fs, description)
    "C2": (256, 7_200_000, 2000, "BASELINE configs[1]: 256-ch ECoG @2 kHz, 60 min session, FULL6 chain"),
    "C1": (128, 1_200_000, 2000, "BASELINE configs[0] shape: 128-ch ECoG @2 kHz, 10 min, FULL6 chain"),
    "tiny": (16, 240_000, 2000, "16-ch @2 kHz, 2 min (debug)"),
    # one recording channel-sharded over the ranks: CAR = colsum -> NCCL all_reduce -> apply (strong scaling)
    "C4": (1024, 10_800_000, 3000, "BASELINE configs[3]: 1024-ch uECoG @3 kHz, 60 min, channel-sharded, NCCL CAR allreduce"),
    "C4small": (64, 1_080_000, 3000, "64-ch @3 kHz, 6 min, channel-sharded (debug)"),
}
SHARDED = {"C4", "C4small"}
# algorithmic bytes per RAW channel-sample, unfused contract of SURVEY.md section 8(d)
FULL6_BYTES_PER_SAMPLE = 39.2
HILBERT_BYTES_PER_SAMPLE = 8.0
# per executed group, same contract (bytes per RAW channel-sample at 2 kHz -> 400 Hz): external inputs
# read once + outputs written once; a fused group counts once; a statistics pre-pass that cannot be
# fused counts one extra read (the CAR column-sum pass in front of the Hilbert kernel).
STEP_BYTES_PER_SAMPLE = {
    "frequency_filter[butter_bandstop]": 8.0, "car_rereference": 8.0, "frequency_filter[butter_bandpass]": 8.0,
    "frequency_filter[hilbert]": 8.0, "downsample": 4.8, "channel_zscore": 2.4,
    "frequency_filter[butter_bandstop]+frequency_filter[butter_bandpass]": 8.0,
    "car_rereference+frequency_filter[hilbert]": 12.0,
}
HILBERT_KEYS = ("car_rereference+frequency_filter[hilbert]", "frequency_filter[hilbert]")
KERNEL_SOURCES = {"hilbert_env8_kernel": "decode_tonal_langauge_b200/csrc/hilbert.cu",
                  "sos_warm_kernel": "decode_tonal_langauge_b200/csrc/sos_common.cuh",
                  "sos_warm_tma_kernel": "decode_tonal_langauge_b200/csrc/sosfilt_tma.cu",
                  "sos_pair_ws_kernel": "decode_tonal_langauge_b200/csrc/sosfilt_pairws.cu",
                  "halfband2_decimate_kernel": "decode_tonal_langauge_b200/csrc/firdecim.cu"}


def source_sha(rel: str) -> str:
    """sha256 of a kernel source with comments and white space removed: a capture stays valid across comment
    edits and goes stale with the first change of code."""
    import hashlib
    import re
    with open(os.path.join(ROOT, rel), "r", encoding="utf-8") as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    return hashlib.sha256(re.sub(r"\s+", " ", src).strip().encode()).hexdigest()[:16]


def ncu_record(kernel: str):
    """Per-kernel ncu facts (DRAM bytes and warp instructions per channel-sample) from the newest
    committed capture -- profiles/r02_traffic.json, written by scripts/ncu_traffic.py from an
    `ncu --set full` report together with the sha256 of the kernel's source file.  A capture taken
    from an older version of the source is STALE and is not reported (None)."""
    for name in ("r02_traffic.json",):
        path = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(path):
            continue
        with open(path) as f:
            e = json.load(f).get(kernel)
        if not e:
            continue
        if e.get("source_sha") != source_sha(KERNEL_SOURCES[kernel]):
            return {"stale": True, "source": e.get("source")}
        return e
    return None


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index),
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------- CPU arms
def _oracle_full6(x, fs):
    from oracle import chains as ochains
    from decode_tonal_langauge_b200.chains import FULL6_STEPS
    return ochains.run_chain(x, fs, FULL6_STEPS)


_WORKER_DATA = {}


def _cpu_worker(args):
    """One process of the reference arm: FULL6 (oracle port) on a block of synthetic channels
    (decode_tonal_langauge_b200.synth, the generator of SURVEY.md section 8d), CAR over the block.
    The block is generated once per process, outside the timed call."""
    block, ch, T, fs, prepare = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    key = (block, ch, T, fs)
    if key not in _WORKER_DATA:
        from decode_tonal_langauge_b200 import synth
        _WORKER_DATA.clear()
        _WORKER_DATA[key] = synth.session_channels(0, range(block * ch, (block + 1) * ch), T, fs, 256)
    if prepare:
        return 0.0, 0.0
    t0 = time.perf_counter()
    y, _ = _oracle_full6(_WORKER_DATA[key], fs)
    return time.perf_counter() - t0, float(np.nanmax(np.abs(y)))


def oracle_rows_of_session(x_dev, rows, fs, col_mean):
    """FULL6 of the oracle for `rows` of the device-resident session the GPU arm timed.  The chain's
    CAR is linear, the same combination of rows at every sample, and every step in front of it is a
    per-row linear filter, so chain(x)[c] = chain_without_car(x[c] - mean_k x[k]) (`col_mean`, float64,
    over ALL channels of the recording).  Returns (float64 oracle, long-double-notch truth, seconds of
    the float64 oracle run)."""
    import torch
    from oracle import chains as ochains
    from oracle import steps as osteps
    from decode_tonal_langauge_b200 import design as D
    from decode_tonal_langauge_b200.chains import FULL6_STEPS
    xin = x_dev[rows].to(torch.float64).cpu().numpy() - col_mean[None]
    no_car = [FULL6_STEPS[0]] + FULL6_STEPS[2:]
    t0 = time.perf_counter()
    ref, _ = ochains.run_chain(xin, fs, no_car)
    dt = time.perf_counter() - t0
    d = D.butter_design([58, 62], fs, 4, False, "bandstop")
    notch_ld = np.asarray(osteps.filtfilt_pad(d.b, d.a, xin, dtype=np.longdouble, reference_edges=True), dtype=np.float64)
    truth, _ = ochains.run_chain(notch_ld, fs, no_car[1:])
    return ref, truth, dt


def max_rel(y, ref):
    y, ref = np.asarray(y, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return float(np.max(np.max(np.abs(y - ref), axis=1) / np.max(np.abs(ref), axis=1)))


def parity_record(got, ref, truth, rows, what):
    """SURVEY.md section 8c rule: float outputs within 1e-5 (max_t|y - ref| / max_t|ref| per channel) of
    the reference; designs with pole radius > 0.99 (the 58-62 Hz notch) are judged against the
    long-double evaluation of the reference's own algorithm: err(gpu, truth) <= max(1e-5, err(ref, truth))."""
    e_gpu, e_ref, e_direct = max_rel(got, truth), max_rel(ref, truth), max_rel(got, ref)
    return {"max_rel": e_gpu, "tol": 1e-5, "pass": bool(e_gpu <= max(1e-5, e_ref)),
            "rule": "err(gpu, long-double truth) <= max(tol, err(float64 reference, truth)); "
                    "max_t|y-ref|/max_t|ref| per channel, worst channel",
            "reference_vs_truth": e_ref, "gpu_vs_float64_reference": e_direct, "rows": list(rows), "what": what}


def run_reference_arm(args):
    """--impl reference: the oracle port of the reference's CPU path on all usable host cores
    (the reference tree is pure Python and does not travel to the GPU box; `kind: port`)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    try:
        import psutil
        avail = psutil.virtual_memory().available
    except Exception:
        avail = 64 << 30
    C, T, fs, desc = WORKLOADS[args.workload]
    ch_per_worker = 2
    per_worker_bytes = T * (8 * 8 * 2 + 16 * 4) * 1.3          # Hilbert temporaries dominate
    workers = int(max(1, min(os.cpu_count() or 1, avail * 0.6 // per_worker_bytes)))
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(workers) as pool:
        pool.map(_cpu_worker, [(w, ch_per_worker, T, fs, True) for w in range(workers)], chunksize=1)
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            pool.map(_cpu_worker, [(w, ch_per_worker, T, fs, False) for w in range(workers)], chunksize=1)
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                times.append(dt)
    total = sum(times)
    units = workers * ch_per_worker * T * len(times)
    value = units / total
    sample = (f"{workers} processes x {ch_per_worker} ch x {T} samples per step = {workers * ch_per_worker} of the {C} "
              f"channels of the workload (bounded sample, full session length); synthetic session of "
              f"decode_tonal_langauge_b200.synth (pink noise + common mode + line noise); oracle port of the reference "
              f"(numpy/scipy, float64), CAR over each {ch_per_worker}-channel block; NOT like for like with the GPU arm: "
              f"fewer channels, CAR over blocks, kind=port")
    line = {"impl": "reference", "metric": "channel_samples_per_sec", "value": value, "unit": "channel-samples/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": desc, "chain": "FULL6", "sample": sample},
            "cpu_baseline": {"value": value, "unit": "channel-samples/s", "cores": workers, "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": "channel-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def _timed_steps(step, steps, barrier, torch):
    """K timed steps bracketed by barrier + synchronize; returns (elapsed ms of this rank, per-step profiles)."""
    profiles = []
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(steps):
        prof = []
        step(prof)
        profiles.append(prof)
    ev1.record()
    barrier()
    return ev0.elapsed_time(ev1), profiles


def _mean_ms(profiles):
    acc = {}
    for prof in profiles:
        for name, a, b in prof:
            acc.setdefault(name, []).append(a.elapsed_time(b))
    return {k: float(np.mean(v)) for k, v in acc.items()}


def host_copy_ceiling(torch, n_in_bytes, out_shape, out_dtype, chunks, steps=3):
    """Pinned H2D of the input and D2H of the result with NO kernels in between, on two copy streams
    (the D2H of chunk i overlaps the H2D of chunk i+1, as far as the host allows): the floor the host
    memory system / PCIe puts under the end-to-end number."""
    dev = torch.device("cuda", torch.cuda.current_device())
    hin = torch.empty(n_in_bytes // 4, dtype=torch.float32, pin_memory=True)
    din = torch.empty(n_in_bytes // 4, dtype=torch.float32, device=dev)
    dout = torch.zeros(out_shape, dtype=out_dtype, device=dev)
    hout = torch.empty(out_shape, dtype=out_dtype, pin_memory=True)
    up, down = torch.cuda.Stream(), torch.cuda.Stream()
    n = hin.numel()
    cut = [n * i // chunks for i in range(chunks + 1)]
    rows = [out_shape[0] * i // chunks for i in range(chunks + 1)]
    times = []
    for it in range(steps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(up):
            for i in range(chunks):
                din[cut[i]:cut[i + 1]].copy_(hin[cut[i]:cut[i + 1]], non_blocking=True)
        up.synchronize()                       # CAR sees every channel before the first result exists
        with torch.cuda.stream(down):
            for i in range(chunks):
                hout[rows[i]:rows[i + 1]].copy_(dout[rows[i]:rows[i + 1]], non_blocking=True)
        down.synchronize()
        if it:
            times.append(time.perf_counter() - t0)
    del hin, din, dout, hout
    return float(np.mean(times))


def run_sharded_record(args, torch, dist, world, rank, local, workload="C4"):
    """One recording channel-sharded over the ranks (BASELINE configs[3]): FULL6 with the CAR column
    sums all-reduced over NCCL.  Returns the sub-record on rank 0 (None elsewhere)."""
    from decode_tonal_langauge_b200 import distributed as D
    from decode_tonal_langauge_b200 import ops, synth
    from decode_tonal_langauge_b200 import _native as nat
    from decode_tonal_langauge_b200.chains import FULL6_STEPS
    C, T, fs, desc = WORKLOADS[workload]
    c_lo, c_hi = D.shard_bounds(C, rank, world)
    x = synth.device_session(c_hi - c_lo, T, fs, seed=0, channel_seed=rank)
    torch.cuda.synchronize()

    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()

    y = None
    steps = max(3, min(args.steps, 10))
    for _ in range(max(3, args.warmup)):
        y, _, _ = D.preprocess_signal_sharded(x, FULL6_STEPS, Namespace(signal_freq=fs), c_lo, C)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    def timed(overlap):
        ar_ev, grp_ev = [], []
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(steps):
            out, _, _ = D.preprocess_signal_sharded(x, FULL6_STEPS, Namespace(signal_freq=fs), c_lo, C, timing=ar_ev,
                                                    profile=grp_ev, overlap_allreduce=overlap)
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / steps, sum(a.elapsed_time(b) for a, b in ar_ev) / steps],
                         dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return out, float(t[0].item()), float(t[1].item()), len(ar_ev) // steps, _mean_ms([grp_ev])

    l0 = nat.launch_count()
    y, ms_seq, ar_seq, _, step_ms = timed(False)           # default: one all-reduce of the whole vector on the compute stream
    launches = nat.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    _, ms_ov, ar_ov, n_ar, step_ms_ov = timed(True)        # option: all-reduce by time-tile groups on a side stream
    ms = torch.tensor([ms_seq], dtype=torch.float64, device="cuda")
    ar = torch.tensor([ar_seq], dtype=torch.float64, device="cuda")
    # parity: three local rows of rank 0 against the oracle, with the GLOBAL column mean (float64, all-reduced)
    col = x.sum(dim=0, dtype=torch.float64)
    dist.all_reduce(col, op=dist.ReduceOp.SUM)
    parity = None
    if rank == 0 and not args.no_cpu:
        rows = [0, (c_hi - c_lo) // 2, c_hi - c_lo - 1]
        ref, truth, _ = oracle_rows_of_session(x, rows, fs, (col / C).cpu().numpy())
        parity = parity_record(y[rows].cpu().numpy(), ref, truth, rows,
                               f"rank 0 rows of the channel-sharded FULL6 output vs the oracle with the global "
                               f"column mean ({C} channels over {world} ranks)")
    barrier()
    del x, y
    ops.release_workspaces()
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    return {"workload": desc, "channels": C, "samples": T, "fs": fs, "scaling": "strong", "n_gpus": world,
            "channels_per_rank": c_hi - c_lo, "steps": steps, "ms_per_step": float(ms.item()),
            "value": C * T / (float(ms.item()) * 1e-3), "unit": "channel-samples/s",
            "allreduce": {"ms": float(ar.item()), "bytes": 4 * T, "share_of_step": float(ar.item()) / ms_seq,
                          "what": "dist.all_reduce(SUM) of the T float32 CAR column sums as ONE collective on the compute "
                                  "stream (the default), CUDA events around the call, max over ranks; it sits between "
                                  "ecog_car_colsum and the Hilbert kernel that subtracts the mean in its load",
                          "overlapped": {"ms_per_step": ms_ov, "collectives_per_step": n_ar, "sum_ms": ar_ov,
                                         "what": "option ECOG_OVERLAP_ALLREDUCE=1: column sums, all-reduce and Hilbert blocks by "
                                                 "time-tile groups, the collectives on a side stream (distributed.OVERLAP_GROUPS); "
                                                 "sum_ms = time the side stream spent in the collectives (waiting for peers included)"}},
            "step_ms": step_ms, "step_ms_overlapped": step_ms_ov, "gpu_launches": int(launches), "clocks": clocks,
            "parity": parity}


def run_c5_record(args, torch, dist, world, rank, local, n_sessions=64):
    """BASELINE configs[4]: a batch of 64 sessions (256 ch x 30 min @ 2 kHz each, 3.7 GB float32) session-sharded
    over the ranks (round robin, no collective), every session streamed file -> pinned -> device -> FULL6 ->
    pinned -> sink through decode_tonal_langauge_b200.sessions.preprocess_sessions.  One session file per rank
    lives in /dev/shm (the reference's block format, np.savez) and is read once per assigned session: the
    whole I/O path runs 64 times without 236 GB of storage.  Whole-batch wall time, max over ranks."""
    from decode_tonal_langauge_b200 import distributed as D
    from decode_tonal_langauge_b200 import ops, sessions, synth
    from decode_tonal_langauge_b200 import runtime as rt
    from decode_tonal_langauge_b200.chains import FULL6_STEPS
    C, T, fs = 256, 3_600_000, 2000
    mine = D.assign_sessions(n_sessions, rank, world)
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    import tempfile
    tmpdir = tempfile.mkdtemp(prefix="ecog_c5_", dir=base)
    path = os.path.join(tmpdir, f"B{rank}_ecog.npz")
    rec = None
    try:
        x = synth.device_session(C, T, fs, seed=100 + rank)
        np.savez(path, data=x.cpu().numpy(), sf=np.float64(fs))
        del x
        torch.cuda.empty_cache()
        sink = lambda i, y, f: float(y[0, :8].sum())
        sessions.preprocess_sessions([path], FULL6_STEPS, sink)                  # warm the pinned pools and plans
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        rt.reset_counters()
        tm = {}
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        t0 = time.perf_counter()
        out = sessions.preprocess_sessions([path] * len(mine), FULL6_STEPS, sink, depth=2, timing=tm)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0, tm.get("read_s", 0.0)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        clocks = sampler.stop() if rank == 0 else None
        ok = len(out) == len(mine) and all(np.isfinite(v) for v in out)
        rec = {"workload": "BASELINE configs[4]: 64-session batch (256 ch, 30 min @ 2 kHz each), session-sharded, FULL6",
               "sessions": n_sessions, "sessions_per_rank": len(mine), "n_gpus": world, "scaling": "strong",
               "s_total": float(dt[0].item()), "value": n_sessions * C * T / float(dt[0].item()),
               "unit": "channel-samples/s", "reader_s_max": float(dt[1].item()),
               "h2d_bytes": int(rt.h2d_bytes), "d2h_bytes": int(rt.d2h_bytes), "ok": bool(ok), "clocks": clocks,
               "path": "np.savez block in /dev/shm -> 4 reader threads readinto() a pinned buffer -> H2D -> FULL6 -> "
                       "float64 D2H into a pinned buffer -> sink (checksum); two buffers per direction, reads, copies "
                       "and kernels of consecutive sessions overlap",
               "note": "every rank reads its own session file once per assigned session (same content, full I/O path)"}
    except Exception as exc:      # noqa: BLE001 -- an optional record must not take the headline line down
        rec = {"unavailable": f"{type(exc).__name__}: {exc}"}
    finally:
        import shutil
        shutil.rmtree(tmpdir, ignore_errors=True)
        ops.release_workspaces()
        torch.cuda.empty_cache()
    return rec if rank == 0 else None


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs CUDA devices; there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    host_cpus = None
    if os.environ.get("ECOG_NO_AFFINITY") is None:
        from decode_tonal_langauge_b200 import runtime as _rt
        host_cpus = _rt.bind_host_to_gpu(local)       # pinned staging buffers next to this GPU's PCIe root
    if world > 1:
        # NCCL prints its version banner on stdout when the communicator is created; the contract
        # is ONE JSON line on stdout, so the banner goes to stderr
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    from decode_tonal_langauge_b200 import _native as nat
    from decode_tonal_langauge_b200 import ops
    from decode_tonal_langauge_b200 import runtime as rt
    from decode_tonal_langauge_b200 import synth
    from decode_tonal_langauge_b200.chains import FULL6_STEPS
    from decode_tonal_langauge_b200.preprocessor import preprocess_signal

    if args.workload in SHARDED:
        if world < 2:
            raise SystemExit(f"--workload {args.workload} is channel-sharded: run it under torchrun with >= 2 ranks")
        rec = run_sharded_record(args, torch, dist, world, rank, local, args.workload)
        if rank == 0:
            line = {"metric": "channel_samples_per_sec", "value": rec["value"], "unit": rec["unit"], "n_gpus": world,
                    "steps": rec["steps"], "warmup": max(3, args.warmup), "ms_per_step": rec["ms_per_step"],
                    "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                    "config": {"workload": rec["workload"], "chain": "FULL6"}, "clocks": rec["clocks"],
                    "gpu_launches": rec["gpu_launches"], "sharded": rec}
            print(json.dumps(line), flush=True)
        dist.destroy_process_group()
        return

    C, T, fs, desc = WORKLOADS[args.workload]
    x = synth.device_session(C, T, fs, seed=rank)              # one session per rank, resident in HBM
    torch.cuda.synchronize()

    def step(profile=None):
        y, f = preprocess_signal(x, FULL6_STEPS, Namespace(signal_freq=fs), profile=profile)
        return y

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        y = step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = nat.launch_count()
    keep = {}

    def timed(prof):
        keep["y"] = step(prof)

    my_ms, profiles = _timed_steps(timed, args.steps, barrier, torch)
    launches = nat.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    y = keep.pop("y")
    ms = torch.tensor([my_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    out_shape = tuple(y.shape)

    # per-group device times (rank 0), dominant kernel = the single-launch Hilbert group
    step_ms = _mean_ms(profiles)
    peak, peak_src = measured_peaks()
    hil_key = next((k for k in HILBERT_KEYS if k in step_ms), None)
    hil_ms = step_ms.get(hil_key)
    chain_gbs = FULL6_BYTES_PER_SAMPLE * C * T / (total_ms / args.steps * 1e-3) / 1e9
    step_roofline = {k: {"ms": v, "alg_gb": STEP_BYTES_PER_SAMPLE[k] * C * T / 1e9,
                         "achieved_gbs": STEP_BYTES_PER_SAMPLE[k] * C * T / (v * 1e-3) / 1e9,
                         "frac": STEP_BYTES_PER_SAMPLE[k] * C * T / (v * 1e-3) / 1e9 / peak}
                     for k, v in step_ms.items() if k in STEP_BYTES_PER_SAMPLE}

    # ---- in-run parity: rows of THIS run's output against the oracle on the same session (rank 0, N = 1)
    cpu, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu:
        rows = [0, C // 2 + 1, C - 1]
        col_mean = x.mean(dim=0, dtype=torch.float64).cpu().numpy()
        ref, truth, dt_cpu = oracle_rows_of_session(x, rows, fs, col_mean)
        parity = parity_record(y[rows].cpu().numpy(), ref, truth, rows,
                               "rows of the timed run's own output (default plans, fused groups) vs the oracle on the "
                               "same synthetic session")
        cpu = {"value": len(rows) * T / dt_cpu, "unit": "channel-samples/s", "cores": 1, "kind": "port",
               "sample": f"oracle FULL6 on {len(rows)} rows x {T} samples of the session the GPU arm timed (full session "
                         f"length; CAR by linearity: the {C}-channel column mean is subtracted up front), 1 thread "
                         f"(the reference's own execution model), {dt_cpu:.1f} s"}
        del ref, truth

    # ---- end to end through the plug-in call with HOST buffers, H2D + D2H inside the timed region
    del y
    e2e = None
    if not args.no_e2e:
        def e2e_run(xin, n, **kw):
            yh, _ = preprocess_signal(xin, FULL6_STEPS, Namespace(signal_freq=fs), **kw)        # warm the pinned pools
            del yh
            rt.reset_counters()
            barrier()
            t0 = time.perf_counter()
            dtype = None
            for _ in range(n):
                yh, _ = preprocess_signal(xin, FULL6_STEPS, Namespace(signal_freq=fs), **kw)
                _ = float(yh[0, :8].sum())
                dtype = str(yh.dtype)
                del yh                      # hand the pinned result buffer back before the next step
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            return float(dt.item()) / n, dtype, rt.h2d_bytes // n, rt.d2h_bytes // n

        host_in = torch.empty((C, T), dtype=torch.float32, pin_memory=True)
        host_in.copy_(x)
        torch.cuda.synchronize()
        xin = host_in.numpy()
        e2e_steps = max(2, min(args.steps, 5))
        sec, out_dtype, h2d, d2h = e2e_run(xin, e2e_steps)
        e2e = {"value": world * C * T / sec, "unit": "channel-samples/s",
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": 1e3 * sec,
               "api": "preprocess_signal(numpy (C,T) float32, pinned) -> numpy float64 (the reference's dtype after its "
                      "first filtfilt); channel-chunked copies overlapped with the kernels",
               "out_dtype": out_dtype, "host_cpus": len(host_cpus) if host_cpus else None}
        # where one end-to-end step spends its time (one extra, untimed call with the trace on)
        rt.trace = []
        barrier()
        yh, _ = preprocess_signal(xin, FULL6_STEPS, Namespace(signal_freq=fs))
        del yh
        tr, rt.trace = rt.trace, None
        ev0 = next(e for l, e, _ in tr if l == "start")
        e2e["breakdown_ms"] = {l: {"device": (ev0.elapsed_time(e) if e is not None else None), "host": 1e3 * h}
                               for l, e, h in tr if l != "start"}
        # the same call with the device's storage dtype handed back (half the return traffic; explicit option)
        sec32, dt32, _, d2h32 = e2e_run(xin, 2, output_dtype=np.float32)
        ceiling = host_copy_ceiling(torch, C * T * 4, out_shape, torch.float64, 12)
        ceil_t = torch.tensor([ceiling], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(ceil_t, op=dist.ReduceOp.MAX)
        e2e["variants"] = {
            "float32_out": {"ms_per_step": 1e3 * sec32, "value": world * C * T / sec32, "out_dtype": dt32,
                            "d2h_bytes_per_step": d2h32, "api": "preprocess_signal(..., output_dtype=np.float32)"},
            "host_copy_ceiling": {"ms_per_step": 1e3 * float(ceil_t.item()), "value": world * C * T / float(ceil_t.item()),
                                  "frac_of_ceiling": float(ceil_t.item()) / sec,
                                  "what": "pinned H2D of the input, then D2H of a float64 result of the same shape, no "
                                          "kernels, all ranks at once (max over ranks): what the host memory system and "
                                          "PCIe allow this contract"}}
        if world == 1:
            pageable = np.empty((C, T), dtype=np.float32)          # what np.load hands a caller
            pageable[:] = xin
            secp, _, _, _ = e2e_run(pageable, 2)
            e2e["variants"]["pageable_in"] = {"ms_per_step": 1e3 * secp, "value": C * T / secp,
                                              "api": "same call, input in pageable host memory (np.load output)"}
            del pageable
        del host_in, xin

    # ---- N > 1: the channel-sharded recording (BASELINE configs[3]) with the NCCL CAR all-reduce
    del x
    ops.release_workspaces()
    torch.cuda.empty_cache()
    sharded = None
    if world > 1 and not args.no_sharded:
        sharded = run_sharded_record(args, torch, dist, world, rank, local, "C4")
    c5 = None
    if not args.no_c5:
        c5 = run_c5_record(args, torch, dist, world, rank, local)

    if rank == 0:
        value = world * C * T * args.steps / (total_ms * 1e-3)
        hil_gbs = HILBERT_BYTES_PER_SAMPLE * C * T / (hil_ms * 1e-3) / 1e9 if hil_ms else None
        rec = ncu_record("hilbert_env8_kernel")
        fresh = bool(rec) and not rec.get("stale")
        traffic = rec["dram_bytes_per_channel_sample"] * C * T / 1e9 if fresh else None
        compute = None
        if fresh and hil_ms and clocks and clocks.get("sm_mhz"):
            # warp instructions issued per clock and SM over the live launch time, against the issue
            # ceiling of this kernel's own instruction mix (scripts/micro/fp32_pipes.cu) and the 4.0 the
            # four schedulers of an SM can issue
            ipc = rec["warp_inst_per_channel_sample"] * C * T / (hil_ms * 1e-3) / (148 * clocks["sm_mhz"] * 1e6)
            compute = {"bound": "issue", "achieved": ipc, "unit": "warp-instructions / clock / SM",
                       "peak": rec.get("mix_ceiling_ipc", 4.0), "frac": ipc / rec.get("mix_ceiling_ipc", 4.0),
                       "frac_of_4_ipc": ipc / 4.0, "thread_inst_per_sample": 32 * rec["warp_inst_per_channel_sample"],
                       "floor_ms_at_4_ipc": rec["warp_inst_per_channel_sample"] * C * T / (4.0 * 148 * clocks["sm_mhz"] * 1e6) * 1e3,
                       "time_ms": hil_ms, "time_note": "live CUDA-event time of the CAR + Hilbert group (the 1.0 ms column-sum "
                                                        "pass is inside it, so the kernel's own rate is ~4 % higher)",
                       "peak_source": rec.get("mix_ceiling_source", "4 schedulers x 1 warp instruction per clock"),
                       "source": rec["source"]}
        pair_key = "frequency_filter[butter_bandstop]+frequency_filter[butter_bandpass]"
        rec_pair = ncu_record("sos_pair_ws_kernel") or ncu_record("sos_warm_tma_kernel")
        if pair_key in step_roofline and rec_pair and not rec_pair.get("stale"):
            # one sweep of the pair per capture entry; the group runs two (forward, backward)
            step_roofline[pair_key]["traffic_gb"] = 2 * rec_pair["dram_bytes_per_channel_sample"] * C * T / 1e9
            step_roofline[pair_key]["traffic_source"] = rec_pair["source"]
        line = {
            "metric": "channel_samples_per_sec", "value": value, "unit": "channel-samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "chain": "FULL6", "channels": C, "samples": T, "fs": fs,
                       "sharding": "one session per rank, no data-path collective"
                                   + ("; plus `sharded`: BASELINE configs[3] channel-sharded with the NCCL CAR all-reduce"
                                      if sharded else ""),
                       "l2": "inputs larger than L2 (7.4 GB per session); no flush needed",
                       "output": list(out_shape), "state_dtype": "f64 IIR state / statistics, f32 storage",
                       "groups": list(step_ms)},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "parity": parity,
            "roofline": {"bound": "issue", "kernel": ("hilbert_env8_kernel (dominant: %.0f%% of the step)" %
                                                      (100 * hil_ms / (total_ms / args.steps))) if hil_ms else None,
                         "achieved": hil_gbs, "peak": peak, "unit": "GB/s",
                         "frac": hil_gbs / peak if hil_gbs else None,
                         "traffic": traffic, "traffic_unit": "GB per launch",
                         "traffic_source": (rec["source"] if fresh else
                                            ("stale: the committed ncu capture predates the kernel source" if rec else None)),
                         "algorithmic_gb_per_launch": HILBERT_BYTES_PER_SAMPLE * C * T / 1e9,
                         "peak_source": peak_src, "compute": compute,
                         "note": "achieved/peak/frac are the HBM roofline of the contract (algorithmic 8 B per sample over the "
                                 "live launch time); the kernel is instruction-issue / shared-memory-pipe / FP32 bound "
                                 "(8.5 FFTs of 4096 points per 3422 samples), see `compute` and DESIGN.md section 3",
                         "chain_achieved": chain_gbs, "chain_frac": chain_gbs / peak,
                         "chain_bytes_per_sample": FULL6_BYTES_PER_SAMPLE, "steps": step_roofline},
            "cpu_baseline": cpu, "step_ms": step_ms, "sharded": sharded, "sessions_c5": c5,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_c3(args):
    """BASELINE configs[2]: ERP epoch extraction + discriminative / active channel selection,
    256 ch x 20k events x 400 samples, source = one 60-min 400 Hz session resident in HBM.
    One step = gather(ERP) + gather(rest) + ANOVA(tone, 4 groups) + ANOVA(syllable, 2) +
    ANOVA(rest vs ERP) + the three run-length selections.  Not the driver's default line."""
    import torch
    from decode_tonal_langauge_b200 import _native as nat
    from decode_tonal_langauge_b200 import ops, selection
    torch.cuda.set_device(0)
    C, T, sf, N, L = 256, 1_440_000, 400, 20_000, 400
    g = torch.Generator(device="cuda"); g.manual_seed(7)
    src = torch.randn((C, T), generator=g, device="cuda", dtype=torch.float32)
    rng = np.random.default_rng(7)
    onsets = np.sort(rng.choice(np.arange(300, 10 * (T // sf - 5)), N, replace=False)) / 10.0
    starts = np.array([int(s * sf) for s in onsets], dtype=np.int64)
    rest_starts = np.arange(0, int(25.0 * sf) - L + 1, L, dtype=np.int64)
    tone = rng.integers(0, 4, N).astype(np.int64)
    syl = rng.integers(0, 2, N).astype(np.int8)
    params = {"p_threshold": 0.01, "active_time_threshold": 0.1}
    peak, peak_src = measured_peaks()

    def step(prof=None):
        def timed(name, fn):
            if prof is None:
                return fn()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); r = fn(); e1.record()
            prof.append((name, e0, e1))
            return r
        ep = timed("epoch_gather", lambda: ops.epoch_gather(src, starts, L))
        rest = timed("epoch_gather_rest", lambda: ops.epoch_gather(src, rest_starts, L))
        data = {"ecog": ep, "ecog_rest": rest, "ecog_sf": np.int64(sf), "tone": tone, "syllable": syl}
        out = [timed("discriminative[tone]", lambda: selection.discriminative_run(data, dict(params, target="tone"))),
               timed("discriminative[syllable]", lambda: selection.discriminative_run(data, dict(params, target="syllable"))),
               timed("active", lambda: selection.active_run(data, params))]
        return out

    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        for _ in range(args.warmup):
            step()
        torch.cuda.synchronize()
        l0 = nat.launch_count()
        profs = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            pr = []
            step(pr)
            profs.append(pr)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    step_ms = {}
    for pr in profs:
        for name, a, b in pr:
            step_ms.setdefault(name, []).append(a.elapsed_time(b))
    step_ms = {k: float(np.mean(v)) for k, v in step_ms.items()}
    elems = N * C * L
    alg = {"epoch_gather": 8.0 * elems, "discriminative[tone]": 4.0 * elems + 16.0 * C * L,
           "discriminative[syllable]": 4.0 * elems + 16.0 * C * L, "active": 4.0 * (elems + len(rest_starts) * C * L) + 16.0 * C * L}
    roof = {k: {"ms": step_ms[k], "alg_gb": v / 1e9, "achieved_gbs": v / (step_ms[k] * 1e-3) / 1e9,
                "frac": v / (step_ms[k] * 1e-3) / 1e9 / peak} for k, v in alg.items()}
    line = {"metric": "epoch_elements_per_sec", "value": elems / (ms * 1e-3), "unit": "epoch channel-samples/s",
            "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32 gather (bit copy) / f64 statistics", "data": "synthetic",
            "config": {"workload": "BASELINE configs[2]: ERP epoch extraction + discriminative channel selection, "
                                   "256 ch x 20k events", "events": N, "channels": C, "epoch_samples": L,
                       "l2": "epoch tensor 8.2 GB >> L2"},
            "gpu_launches": int(nat.launch_count() - l0),
            "roofline": {"bound": "hbm", "kernel": "anova_f_kernel (discriminative[tone])", "peak": peak, "unit": "GB/s",
                         "achieved": roof["discriminative[tone]"]["achieved_gbs"], "frac": roof["discriminative[tone]"]["frac"],
                         "traffic": None, "peak_source": peak_src, "steps": roof},
            "step_ms": step_ms}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS) + ["C3"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the channel-sharded configs[3] sub-record")
    ap.add_argument("--no-c5", action="store_true", help="skip the 64-session batch (configs[4]) sub-record")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        print("note: fewer than 3 warm-up steps; numbers are not reportable", file=sys.stderr)
    if args.impl == "reference":
        run_reference_arm(args)
    elif args.workload == "C3":
        run_c3(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
