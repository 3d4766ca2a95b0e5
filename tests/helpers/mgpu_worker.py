"""torchrun worker for tests/test_gpu_multi.py: channel-sharded FULL6 + selection over NCCL.

Every rank generates the same small session, keeps its contiguous block of channels on its own
GPU, runs ``preprocess_signal_sharded`` (CAR = ecog_car_colsum -> all_reduce(SUM) -> ecog_car_apply),
gathers, and rank 0 compares with the single-GPU run of the whole array on its device."""
import os
import sys
from argparse import Namespace

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    from decode_tonal_langauge_b200 import distributed as D
    from decode_tonal_langauge_b200 import ops, synth
    from decode_tonal_langauge_b200.chains import FULL6_STEPS
    from decode_tonal_langauge_b200.preprocessor import preprocess_signal

    fs, C, T = 2000, 22, 60000
    x, _ = synth.session(9, C, T, fs, n_events=0)
    excl = [1, 20]
    steps = [dict(s) for s in FULL6_STEPS]
    steps[1] = {"module": "preprocess.car_rereference", "params": {"exclude_channels": excl}}
    lo, hi = D.shard_bounds(C, rank, world)
    xl = torch.from_numpy(x[lo:hi].copy()).cuda()
    yl, f, bands = D.preprocess_signal_sharded(xl, steps, Namespace(signal_freq=fs), lo, C)
    full = D.gather_channels(yl, C, bands)
    # the optional overlapped form: column sums / all-reduce / Hilbert blocks by time-tile groups on a side stream
    yo, fo, _ = D.preprocess_signal_sharded(xl, steps, Namespace(signal_freq=fs), lo, C, overlap_allreduce=True)
    same = float(((yo - yl).abs().amax(dim=1) / yl.abs().amax(dim=1)).max())
    runs = (torch.arange(lo, hi, device="cuda", dtype=torch.int32) * 7) % 50
    sel = D.gather_selection(runs, lo, C, 40)
    ok = True
    if rank == 0:
        ref, f0 = preprocess_signal(torch.from_numpy(x).cuda(), steps, Namespace(signal_freq=fs))
        err = float(((full - ref).abs().amax(dim=1) / ref.abs().amax(dim=1)).max())
        want = [int(c) for c in np.nonzero((np.arange(C) * 7) % 50 > 40)[0]]
        ok = f == f0 == fo == 400 and full.shape == ref.shape and err < 5e-6 and sel == want and same < 1e-6
        print(f"mgpu world={world} err={err:.2e} overlapped_vs_single={same:.2e} sel_ok={sel == want} ok={ok}", flush=True)
    if same >= 1e-6:
        ok = False
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
