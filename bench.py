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
# per step, same contract (bytes per RAW channel-sample at 2 kHz -> 400 Hz)
STEP_BYTES_PER_SAMPLE = {
    "frequency_filter[butter_bandstop]": 8.0, "car_rereference": 8.0, "frequency_filter[butter_bandpass]": 8.0,
    "frequency_filter[hilbert]": 8.0, "downsample": 4.8, "channel_zscore": 2.4,
}


def ncu_traffic(kernel: str, channels: int, samples: int):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture
    (profiles/r01_traffic.json: bytes per channel-sample measured by ncu, scaled to this launch)."""
    path = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        t = json.load(f)
    e = t.get(kernel)
    if not e:
        return None
    return {"gb_per_launch": e["dram_bytes_per_channel_sample"] * channels * samples / 1e9,
            "source": e["source"]}


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


def _cpu_worker(args):
    seed, ch, T, fs = args
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    rng = np.random.default_rng(seed)
    x = (rng.standard_normal((ch, T)) * 30).astype(np.float32)
    t0 = time.perf_counter()
    y, _ = _oracle_full6(x, fs)
    return time.perf_counter() - t0, float(np.nanmax(np.abs(y)))


def cpu_baseline_single(T, fs, channels=3):
    """The reference's own execution model: one thread."""
    dt, _ = _cpu_worker((1234, channels, T, fs))
    return {"value": channels * T / dt, "unit": "channel-samples/s", "cores": 1, "kind": "port",
            "sample": f"oracle FULL6 on {channels} ch x {T} samples (full session length), 1 thread, {dt:.1f} s"}


def run_reference_arm(args):
    """--impl reference: the oracle port of the reference's CPU path on all usable host cores."""
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
        for step in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            pool.map(_cpu_worker, [(step * 1000 + w, ch_per_worker, T, fs) for w in range(workers)])
            dt = time.perf_counter() - t0
            if step >= args.warmup:
                times.append(dt)
    total = sum(times)
    units = workers * ch_per_worker * T * len(times)
    value = units / total
    sample = (f"{workers} processes x {ch_per_worker} ch x {T} samples per step (bounded sample of {C} ch); "
              f"oracle port of the reference (numpy/scipy), CAR over each block")
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
    from decode_tonal_langauge_b200 import runtime as rt
    from decode_tonal_langauge_b200 import synth
    from decode_tonal_langauge_b200.chains import FULL6_STEPS
    from decode_tonal_langauge_b200.preprocessor import preprocess_signal

    C, T, fs, desc = WORKLOADS[args.workload]
    sharded = args.workload in SHARDED
    if sharded:
        from decode_tonal_langauge_b200 import distributed as D
        c_lo, c_hi = D.shard_bounds(C, rank, world)
        # every rank draws the same common-mode / line terms (same seed) and its own channel noise
        x = synth.device_session(c_hi - c_lo, T, fs, seed=0, channel_seed=rank)
        C_local = c_hi - c_lo
    else:
        x = synth.device_session(C, T, fs, seed=rank)              # one session per rank, resident in HBM
        C_local = C
    torch.cuda.synchronize()

    def step(profile=None):
        if sharded:
            y, f, _ = D.preprocess_signal_sharded(x, FULL6_STEPS, Namespace(signal_freq=fs), c_lo, C)
            return y
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
    profiles = []
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        prof = []
        y = step(prof)
        profiles.append(prof)
    ev1.record()
    barrier()
    launches = nat.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    ms = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    total_ms = float(ms.item())
    out_shape = tuple(y.shape)

    # per-step device times (rank 0), dominant kernel = the single-launch Hilbert step
    step_ms = {}
    for prof in profiles:
        for name, a, b in prof:
            step_ms.setdefault(name, []).append(a.elapsed_time(b))
    step_ms = {k: float(np.mean(v)) for k, v in step_ms.items()}
    hil_key = "frequency_filter[hilbert]"
    peak, peak_src = measured_peaks()
    hil_ms = step_ms.get(hil_key)
    hil_gbs = HILBERT_BYTES_PER_SAMPLE * C_local * T / (hil_ms * 1e-3) / 1e9 if hil_ms else None
    chain_gbs = FULL6_BYTES_PER_SAMPLE * C_local * T / (total_ms / args.steps * 1e-3) / 1e9
    step_roofline = {k: {"ms": v, "alg_gb": STEP_BYTES_PER_SAMPLE[k] * C_local * T / 1e9,
                         "achieved_gbs": STEP_BYTES_PER_SAMPLE[k] * C_local * T / (v * 1e-3) / 1e9,
                         "frac": STEP_BYTES_PER_SAMPLE[k] * C_local * T / (v * 1e-3) / 1e9 / peak}
                     for k, v in step_ms.items() if k in STEP_BYTES_PER_SAMPLE}
    traffic = ncu_traffic("hilbert_env8_kernel", C_local, T)

    # ---- end to end through the plug-in call with HOST buffers (pinned), H2D + D2H inside
    del y
    e2e = None
    if not args.no_e2e and not sharded:
        host_in = torch.empty((C, T), dtype=torch.float32, pin_memory=True)
        host_in.copy_(x)
        torch.cuda.synchronize()
        xin = host_in.numpy()
        e2e_steps = max(2, min(args.steps, 5))
        rt.reset_counters()
        yh, _ = preprocess_signal(xin, FULL6_STEPS, Namespace(signal_freq=fs))        # warm the pinned pools
        del yh
        rt.reset_counters()
        barrier()
        t0 = time.perf_counter()
        out_dtype = None
        for _ in range(e2e_steps):
            yh, _ = preprocess_signal(xin, FULL6_STEPS, Namespace(signal_freq=fs))
            checksum = float(yh[0, :8].sum())
            out_dtype = str(yh.dtype)
            del yh                      # hand the pinned result buffer back before the next step
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * C * T * e2e_steps / float(dt.item()), "unit": "channel-samples/s",
               "h2d_bytes_per_step": rt.h2d_bytes // e2e_steps, "d2h_bytes_per_step": rt.d2h_bytes // e2e_steps,
               "steps": e2e_steps, "ms_per_step": 1e3 * float(dt.item()) / e2e_steps,
               "api": "preprocess_signal(numpy (C,T) float32 pinned) -> numpy float64 (reference dtype); "
                      "channel-chunked copies overlapped with the kernels",
               "out_dtype": out_dtype,
               "host_cpus": len(host_cpus) if host_cpus else None}     # cores this rank is pinned to (GPU-local NUMA node)
        del host_in

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline_single(T, fs)

    if rank == 0:
        units = C * T if sharded else world * C * T
        value = units * args.steps / (total_ms * 1e-3)
        line = {
            "metric": "channel_samples_per_sec", "value": value, "unit": "channel-samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if sharded else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "chain": "FULL6", "channels": C, "samples": T, "fs": fs,
                       "sharding": ("channels of one recording split over the ranks; CAR column sums "
                                    "all-reduced over NCCL (T floats per step)") if sharded else
                                   "one session per rank, no data-path collective",
                       "l2": "inputs larger than L2 (7.4 GB per session); no flush needed",
                       "output": list(out_shape), "state_dtype": "f64 IIR state / statistics, f32 storage"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "hilbert_env8_kernel (dominant: %.0f%% of the step)" %
                         (100 * hil_ms / (total_ms / args.steps)) if hil_ms else None,
                         "achieved": hil_gbs, "peak": peak, "unit": "GB/s",
                         "frac": hil_gbs / peak if hil_gbs else None,
                         "traffic": traffic["gb_per_launch"] if traffic else None, "traffic_unit": "GB per launch",
                         "traffic_source": traffic["source"] if traffic else None,
                         "algorithmic_gb_per_launch": HILBERT_BYTES_PER_SAMPLE * C_local * T / 1e9,
                         "peak_source": peak_src,
                         "note": "issue / shared-memory-pipe / FP32 bound kernel (8.5 FFTs of 4096 points per 3422 samples; ncu: "
                                 "issue 68 %, LSU data pipe 63 %, FMA pipe 53 %, DRAM 7.5 %), not HBM bound; "
                                 "see DESIGN.md section 3",
                         "chain_achieved": chain_gbs, "chain_frac": chain_gbs / peak,
                         "chain_bytes_per_sample": FULL6_BYTES_PER_SAMPLE, "steps": step_roofline},
            "cpu_baseline": cpu, "step_ms": step_ms,
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
