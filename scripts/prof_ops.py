"""Run single operators at chosen shapes (profiling helper for ncu / CUDA-event timing).

    python scripts/prof_ops.py hilbert 128 1200000
    python scripts/prof_ops.py resample|resample1 32 7200000      (two-stage | single whole-row FFT)
    python scripts/prof_ops.py notch|bandpass|notch_scan|bandpass_scan|car|zscore|fir C T [reps]
    python scripts/prof_ops.py all 256 7200000                     (every operator, one line each)
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from decode_tonal_langauge_b200 import fftplan as FP  # noqa: E402
from decode_tonal_langauge_b200 import design as DSG  # noqa: E402
from decode_tonal_langauge_b200 import ops  # noqa: E402

op, C, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
fs = float(os.environ.get("ECOG_PROF_FS", "2000"))
x = torch.randn((C, T), device="cuda") * 30


def fir():
    pre = FP.predecimation(T, T // 5)
    if pre.halfband is not None:
        return ops.halfband2_decimate(x, *pre.halfband)
    return ops.fir_decimate(x, pre.taps, pre.offset, pre.D)


def _with_env(key, val, fn):
    old = os.environ.get(key)
    os.environ[key] = val
    try:
        return fn()
    finally:
        if old is None:
            del os.environ[key]
        else:
            os.environ[key] = old


FN = {
    "hilbert": lambda: ops.hilbert(x, fs, [70.0, 150.0]),
    "resample": lambda: ops.fft_resample(x, T // 5),
    "resample1": lambda: ops.fft_resample(x, T // 5, two_stage=False),
    "fir": fir,
    "notch": lambda: ops.butter(x, [58, 62], fs, 4, False, "bandstop"),
    "bandpass": lambda: ops.butter(x, [70, 150], fs, 4, False, "bandpass"),
    "notch_scan": lambda: ops.butter(x, [58, 62], fs, 4, False, "bandstop", mode="scan"),
    "bandpass_scan": lambda: ops.butter(x, [70, 150], fs, 4, False, "bandpass", mode="scan"),
    "car": lambda: ops.car(x),
    "zscore": lambda: ops.zscore(x),
    "pair": lambda: ops.sosfilt_pair(x, DSG.butter_design([58, 62], fs, 4, False, "bandstop"),
                                     DSG.butter_design([70, 150], fs, 4, False, "bandpass")),
    "notch_tma": lambda: ops.butter(x, [58, 62], fs, 4, False, "bandstop", mode="tma"),
    "bandpass_tma": lambda: ops.butter(x, [70, 150], fs, 4, False, "bandpass", mode="tma"),
    "pair_tma": lambda: _with_env("ECOG_SOS_TMA", "1", FN["pair"]),
    "pair256": lambda: _with_env("ECOG_PAIR_TPS", "256", FN["pair"]),
    "pair384": lambda: _with_env("ECOG_PAIR_TPS", "384", FN["pair"]),
    "pair512": lambda: _with_env("ECOG_PAIR_TPS", "512", FN["pair"]),
    "pair_f64": lambda: _with_env("ECOG_PAIR_F32", "0", FN["pair"]),
    "colsum": lambda: ops.car_colsum(x),
    "hilbert_car": lambda: ops.hilbert(x, fs, [70.0, 150.0], car=(COLSUM, C)),
}
COLSUM = torch.zeros(T, device="cuda") if "hilbert_car" in op or op == "all" else None


def run(name):
    fn = FN[name]
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        y = fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    del y
    print(f"{name} C={C} T={T} tps={os.environ.get('ECOG_SOS_TPS', '512')}: {ms:.3f} ms  "
          f"{C * T / ms / 1e6:.1f} G ch-samp/s", flush=True)


# bring the clocks up before the first timed operator (an idle B200 sits at 120 MHz)
_w = torch.randn((4096, 4096), device="cuda")
for _ in range(0 if os.environ.get("ECOG_PROF_NO_SPINUP") else 200):
    _w = (_w @ _w).clamp_(-1, 1)
torch.cuda.synchronize()
del _w

for name in (FN if op == "all" else op.split(",")):
    run(name)
