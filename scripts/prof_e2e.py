"""Where the end-to-end time of preprocess_signal(host numpy) goes (GPU box only)."""
import os, sys, time
from argparse import Namespace
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from decode_tonal_langauge_b200 import runtime as rt
from decode_tonal_langauge_b200.chains import FULL6_STEPS
from decode_tonal_langauge_b200.preprocessor import preprocess_signal

C, T, fs = 256, 7_200_000, 2000
host = torch.empty((C, T), dtype=torch.float32, pin_memory=True)
host.normal_()
xin = host.numpy()
def t(f, n=3):
    torch.cuda.synchronize(); out = []
    for _ in range(n):
        t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); out.append(time.perf_counter() - t0); del r
    return min(out) * 1e3
print("from_numpy pinned?", torch.from_numpy(xin).is_pinned())
print("H2D pinned tensor      %.1f ms" % t(lambda: host.to("cuda", non_blocking=True)))
print("H2D rt.to_device(numpy) %.1f ms" % t(lambda: rt.to_device(xin)))
xd = rt.to_device(xin)
print("device chain           %.1f ms" % t(lambda: preprocess_signal(xd, FULL6_STEPS, Namespace(signal_freq=fs))[0]))
yd, _ = preprocess_signal(xd, FULL6_STEPS, Namespace(signal_freq=fs))
print("D2H rt.to_host f64     %.1f ms" % t(lambda: rt.to_host(yd, np.float64)))
print("D2H rt.to_host f32     %.1f ms" % t(lambda: rt.to_host(yd, np.float32)))
print("pinned alloc 2.9GB     %.1f ms" % t(lambda: torch.empty((C, T // 5), dtype=torch.float64, pin_memory=True), 2))
print("e2e preprocess_signal  %.1f ms" % t(lambda: preprocess_signal(xin, FULL6_STEPS, Namespace(signal_freq=fs))[0]))
