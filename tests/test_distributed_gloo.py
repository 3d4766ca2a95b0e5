"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: channel sharding with the CAR
all-reduce, band-major gather, selection gather, session assignment.  The compute backend is
the numpy oracle here (the product always uses the CUDA ops)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from decode_tonal_langauge_b200 import distributed as D


class NumpyBackend:
    """oracle stand-in for ecog_car_colsum / ecog_car_apply on CPU tensors"""
    @staticmethod
    def car_colsum(x, weights=None):
        w = torch.ones(x.shape[0]) if weights is None else weights
        return (x * w[:, None]).sum(dim=0)

    @staticmethod
    def car_apply(x, colsum, n_included):
        return x - (colsum / n_included)[None, :]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, C, T, excl, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)
        x = rng.standard_normal((C, T)).astype(np.float32)
        lo, hi = D.shard_bounds(C, rank, world)
        y_local = D.car_sharded(torch.from_numpy(x[lo:hi].copy()), lo, C, excl, backend=NumpyBackend())
        full = D.gather_channels(y_local, C, bands=1)
        # band-major layout after a 2-band frequency_filter: local rows [band0 shard; band1 shard]
        two = torch.cat([y_local, -y_local], dim=0)
        full2 = D.gather_channels(two, C, bands=2)
        # CAR over the band-major array of a 2-band frequency_filter: (2 C) rows, exclusions are global rows
        xb = rng.standard_normal((2 * C, T)).astype(np.float32)
        local_b = torch.from_numpy(np.concatenate([xb[lo:hi], xb[C + lo:C + hi]]).copy())
        excl_b = [e for e in (1, C + 2) if e < 2 * C]
        yb = D.car_sharded(local_b, lo, C, excl_b, backend=NumpyBackend(), bands=2)
        full_b = D.gather_channels(yb, C, bands=2)
        runs = torch.arange(lo, hi, dtype=torch.int32) * 7 % 50
        sel = D.gather_selection(runs, lo, C, 40)
        if rank == 0:
            np.savez(os.path.join(out_dir, "out.npz"), full=full.numpy(), full2=full2.numpy(), sel=np.array(sel),
                     full_b=full_b.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("C,excl", [(7, []), (10, [0, 9, 4])])
def test_channel_sharded_car_over_gloo(tmp_path, C, excl):
    from oracle import steps as S
    T, world = 257, 2
    mp.spawn(_worker, args=(world, _free_port(), C, T, excl, str(tmp_path)), nprocs=world, join=True)
    out = np.load(tmp_path / "out.npz")
    x = np.random.default_rng(0).standard_normal((C, T)).astype(np.float32)
    ref = S.car_rereference(x, excl)
    assert np.max(np.abs(out["full"] - ref)) < 1e-6
    assert np.max(np.abs(out["full2"][:C] - ref)) < 1e-6 and np.max(np.abs(out["full2"][C:] + ref)) < 1e-6
    rng = np.random.default_rng(0)
    rng.standard_normal((C, T))
    xb = rng.standard_normal((2 * C, T)).astype(np.float32)
    assert np.max(np.abs(out["full_b"] - S.car_rereference(xb, [1, C + 2]))) < 1e-6
    runs = np.arange(C) * 7 % 50
    assert out["sel"].tolist() == [int(c) for c in np.nonzero(runs > 40)[0]]


def test_shard_bounds_and_sessions():
    for n in (1, 7, 256, 1024):
        for world in (1, 2, 3, 8):
            spans = [D.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1
    assert D.assign_sessions(64, 3, 8) == list(range(3, 64, 8))
    assert sorted(s for r in range(4) for s in D.assign_sessions(10, r, 4)) == list(range(10))


def test_car_sharded_validates_like_reference():
    x = torch.zeros(2, 4)
    with pytest.raises(ValueError):
        D.car_sharded(x, 0, 2, exclude_channels="1", backend=NumpyBackend())
    with pytest.raises(ValueError):
        D.car_sharded(x, 0, 2, exclude_channels=[5], backend=NumpyBackend())


def test_local_exclusions_band_major():
    # 10 channels, 3 bands, shard = channels [4, 7): local rows [b0: 4,5,6 | b1: 4,5,6 | b2: 4,5,6]
    local, n_inc = D.local_exclusions([5, 10 + 4, 20 + 9, 3], 4, 9, 10, bands=3)
    assert local == [1, 3] and n_inc == 30 - 4
    assert D.local_exclusions([], 0, 5, 5) == ([], 5)
    with pytest.raises(ValueError):
        D.local_exclusions([30], 4, 9, 10, bands=3)
