"""Batch runner for many independent recordings (sessions / blocks) on one GPU per process.

ref: preprocess/pipelines/subject_block.py:74-101 loops the blocks of every subject one after the
other -- load (preprocess/io/tdt_blocks.py:6-18), preprocess, save (tdt_blocks.py:21-35) -- with no
state shared between blocks.  Blocks therefore shard over GPUs with no collective
(``distributed.assign_sessions``), and on each GPU the three phases of consecutive blocks overlap:

    reader thread :  file -> pinned host buffer        (block i+1)   readinto(), no intermediate copy
    copy stream   :  pinned -> device                  (block i+1)
    main stream   :  preprocess_signal on the device   (block i)
    copy stream 2 :  device -> pinned result buffer    (block i-1)
    writer thread :  sink(result)                      (block i-1)   e.g. np.savez like the reference

``depth`` pinned buffers per direction bound the host memory (depth x one recording).  The input
files are the reference's own block format: an uncompressed ``.npz`` with ``data`` (C, T) and ``sf``.
"""
from __future__ import annotations

import queue
import struct
import threading
import zipfile
from argparse import Namespace
from dataclasses import dataclass
from typing import Callable, Iterable, List, Optional, Sequence, Tuple

import numpy as np

from . import runtime as rt


@dataclass
class NpzMember:
    path: str
    offset: int          # file offset of the raw array bytes
    shape: Tuple[int, ...]
    dtype: np.dtype
    fortran: bool


def locate_npz_array(path: str, key: str = "data") -> NpzMember:
    """Where the raw bytes of ``key`` live inside an UNCOMPRESSED .npz (np.savez): zip local header +
    .npy header are parsed, nothing is read.  Raises ValueError for compressed members."""
    with zipfile.ZipFile(path) as zf:
        info = zf.getinfo(key + ".npy")
        if info.compress_type != zipfile.ZIP_STORED:
            raise ValueError(f"{path}:{key} is compressed; np.savez (stored) is the reference's format")
        with open(path, "rb") as fh:
            fh.seek(info.header_offset)
            local = fh.read(30)
            if local[:4] != b"PK\x03\x04":
                raise ValueError(f"{path}: bad zip local header")
            n_name, n_extra = struct.unpack("<HH", local[26:30])
            start = info.header_offset + 30 + n_name + n_extra
            fh.seek(start)
            magic = fh.read(8)
            if magic[:6] != b"\x93NUMPY":
                raise ValueError(f"{path}:{key} is not an .npy member")
            major = magic[6]
            hlen = struct.unpack("<H", fh.read(2))[0] if major == 1 else struct.unpack("<I", fh.read(4))[0]
            header = eval(fh.read(hlen).decode("latin1"), {"__builtins__": {}}, {"True": True, "False": False})     # noqa: S307 (npy header literal)
            offset = start + 8 + (2 if major == 1 else 4) + hlen
    return NpzMember(path, offset, tuple(header["shape"]), np.dtype(header["descr"]), bool(header["fortran_order"]))


def read_npz_scalar(path: str, key: str = "sf"):
    with np.load(path) as z:
        return z[key][()]


class _PinnedPool:
    """``depth`` reusable pinned byte buffers (grow-only), handed out through a queue."""

    def __init__(self, depth: int):
        self.free: "queue.Queue" = queue.Queue()
        for _ in range(depth):
            self.free.put(None)

    def get(self, nbytes: int):
        import torch
        buf = self.free.get()
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        return buf

    def put(self, buf) -> None:
        self.free.put(buf)


READ_THREADS = 8          # concurrent readers per file: one thread copies ~3-5 GB/s out of the page cache


def _read_range(path: str, offset: int, view: memoryview) -> None:
    with open(path, "rb", buffering=0) as fh:
        fh.seek(offset)
        done, total = 0, len(view)
        while done < total:
            n = fh.readinto(view[done:min(total, done + (64 << 20))])
            if not n:
                raise IOError(f"{path}: short read ({done} of {total} bytes at offset {offset})")
            done += n


def _read_into(member: NpzMember, dst: np.ndarray, threads: int = READ_THREADS) -> None:
    """File bytes straight into the (pinned) destination: `threads` readers on disjoint byte ranges
    (read() releases the GIL, so they run concurrently), no intermediate buffer."""
    view = memoryview(dst.reshape(-1).view(np.uint8))
    total = len(view)
    if threads <= 1 or total < (64 << 20):
        _read_range(member.path, member.offset, view)
        return
    cuts = [total * k // threads // 4096 * 4096 for k in range(threads)] + [total]
    errs: list = []

    def work(k):
        try:
            _read_range(member.path, member.offset + cuts[k], view[cuts[k]:cuts[k + 1]])
        except BaseException as exc:      # noqa: BLE001
            errs.append(exc)

    ts = [threading.Thread(target=work, args=(k,), daemon=True) for k in range(threads)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    if errs:
        raise errs[0]


def preprocess_sessions(sources: Sequence, steps: List[dict], sink: Optional[Callable] = None, depth: int = 2,
                        output_dtype=None, fuse: Optional[bool] = None, base_params: Optional[Namespace] = None,
                        timing: Optional[dict] = None) -> list:
    """Run ``steps`` over every source, overlapping file reads, PCIe copies, kernels and the sink.

    ``sources``: paths of ``.npz`` blocks (keys ``data``, ``sf``), or ``(array, sf)`` tuples already in host
    memory.  ``sink(index, result (numpy, pinned-backed, valid until the sink returns), signal_freq)`` runs on
    the writer thread; its return values are collected in source order (None without a sink).
    ``timing``: optional dict, filled with per-phase seconds (read, h2d+compute+d2h critical path, sink)."""
    import time

    import torch
    from .preprocessor import preprocess_signal
    from . import steps as S

    dev = rt.device()
    main = torch.cuda.current_stream()
    up, down = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    in_pool, out_pool = _PinnedPool(depth), _PinnedPool(depth)
    loaded: "queue.Queue" = queue.Queue(maxsize=depth)
    finished: "queue.Queue" = queue.Queue(maxsize=depth)
    results = [None] * len(sources)
    errors: list = []
    t_read = [0.0]
    t_sink = [0.0]

    def reader():
        try:
            for i, src in enumerate(sources):
                t0 = time.perf_counter()
                if isinstance(src, (tuple, list)):
                    arr, sf = np.asarray(src[0]), src[1]
                    if arr.ndim != 2:
                        raise ValueError(f"expected a (channels, time) array, got shape {arr.shape}")
                    buf = in_pool.get(arr.nbytes)
                    host = buf[:arr.nbytes].numpy().view(arr.dtype).reshape(arr.shape)
                    np.copyto(host, arr)
                else:
                    m = locate_npz_array(src, "data")
                    if m.fortran or len(m.shape) != 2:
                        raise ValueError(f"{src}: expected a C-ordered (channels, time) array")
                    nbytes = int(np.prod(m.shape)) * m.dtype.itemsize
                    buf = in_pool.get(nbytes)
                    host = buf[:nbytes].numpy().view(m.dtype).reshape(m.shape)
                    _read_into(m, host)
                    sf = read_npz_scalar(src, "sf")
                t_read[0] += time.perf_counter() - t0
                loaded.put((i, buf, host, sf))
        except BaseException as exc:      # noqa: BLE001 -- handed to the caller's thread
            errors.append(exc)
        finally:
            loaded.put(None)

    def writer():
        try:
            while True:
                item = finished.get()
                if item is None:
                    return
                i, buf, host, freq, ev = item
                ev.synchronize()
                t0 = time.perf_counter()
                if sink is not None:
                    results[i] = sink(i, host, freq)
                t_sink[0] += time.perf_counter() - t0
                out_pool.put(buf)
        except BaseException as exc:      # noqa: BLE001
            errors.append(exc)
            while finished.get() is not None:
                pass

    rd, wr = threading.Thread(target=reader, daemon=True), threading.Thread(target=writer, daemon=True)
    rd.start()
    wr.start()
    t_start = time.perf_counter()
    try:
        while True:
            item = loaded.get()
            if item is None or errors:
                break
            i, buf, host, sf = item
            src_t = torch.from_numpy(host)
            with torch.cuda.stream(up):
                x = torch.empty(src_t.shape, dtype=src_t.dtype, device=dev)
                x.copy_(src_t, non_blocking=True)
                arrived = torch.cuda.Event()
                arrived.record(up)
            rt.h2d_bytes += src_t.numel() * src_t.element_size()
            main.wait_event(arrived)
            x.record_stream(main)
            params = Namespace(**vars(base_params)) if base_params is not None else Namespace()
            params.signal_freq = sf
            xf = x if x.dtype == torch.float32 else x.to(torch.float32)
            y, freq = preprocess_signal(xf, steps, params, fuse=fuse)
            # the input buffer may be refilled as soon as its copy has been consumed
            arrived.synchronize()
            in_pool.put(buf)
            ref_dt = np.dtype(host.dtype)
            for step in steps:
                ref_dt = S.reference_dtype_after(step["module"].split(".")[-1], step.get("params", {}) or {}, ref_dt)
            out_dt = np.dtype(output_dtype) if output_dtype is not None else rt.output_dtype(ref_dt)
            td = getattr(torch, out_dt.name)
            yo = y if y.dtype == td else y.to(td)
            done = torch.cuda.Event()
            done.record(main)
            nbytes = yo.numel() * yo.element_size()
            obuf = out_pool.get(nbytes)
            ohost_t = obuf[:nbytes].view(td).view(yo.shape)
            down.wait_event(done)
            with torch.cuda.stream(down):
                ohost_t.copy_(yo, non_blocking=True)
                copied = torch.cuda.Event()
                copied.record(down)
            yo.record_stream(down)
            rt.d2h_bytes += nbytes
            finished.put((i, obuf, ohost_t.numpy(), freq, copied))
            del x, xf, y, yo
    finally:
        finished.put(None)
        wr.join()
        rd.join(timeout=1.0)
    main.synchronize()
    if errors:
        raise errors[0]
    if timing is not None:
        timing.update({"total_s": time.perf_counter() - t_start, "read_s": t_read[0], "sink_s": t_sink[0],
                       "sessions": len(sources)})
    return results


def save_block_sink(setup_dir: str, subject_id: int, block_ids: Iterable[int], modality: str = "ecog") -> Callable:
    """Sink that writes ``<setup>/subject_<id>/B<block>_<modality>.npz`` like the reference
    (ref: preprocess/io/tdt_blocks.py:21-35)."""
    import os
    ids = list(block_ids)
    out_dir = os.path.join(setup_dir, f"subject_{subject_id}")
    os.makedirs(out_dir, exist_ok=True)

    def sink(i, result, freq):
        path = os.path.join(out_dir, f"B{ids[i]}_{modality}.npz")
        np.savez(path, data=result, sf=freq)
        return path

    return sink
