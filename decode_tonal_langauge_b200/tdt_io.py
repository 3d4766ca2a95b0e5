"""Reader for the continuous stream stores of a TDT (Tucker-Davis Technologies) tank block.

ref: preprocess/io/tdt_blocks.py:6-18 calls ``tdt.read_block(block_path)`` and uses nothing but
``blk.streams.<STORE>.data`` (channels x samples) and ``.fs`` of two stream stores (``EOG1``: ECoG,
``ANIN``: audio).  The ``tdt`` wheel is not part of this image and there is no network, so the stream
part of the format is read natively here from the published TTank layout:

``<block>/<tank>_<block>.tsq``   fixed 40-byte event headers, little endian::

    int32   size        record length in 32-bit words, 10 header words included
    int32   type        event type; 0x8101 = stream
    char[4] code        store name ("EOG1", "ANIN", ...); 1 / 2 in the first / last record (block start / stop marks)
    uint16  channel     1-based
    uint16  sort_code
    float64 timestamp   seconds (unix time); start / stop time in the two marks
    int64   fp_loc      byte offset of the payload in the .tev file (a float64 "strobe" for scalar events)
    int32   format      0 float32, 1 int32, 2 int16, 3 int8, 4 float64, 5 int64
    float32 frequency   sampling rate of the store

``<block>/<tank>_<block>.tev``   payloads: (size - 10) 32-bit words per stream header at ``fp_loc``.

A store's chunks are concatenated per channel in time order; channels become rows (lowest channel
first), exactly the ``(n_channels, n_samples)`` array ``tdt.read_block`` hands back.  Snippets, epocs,
scalars and SEV side files are skipped (the reference does not read them).  ``write_block`` produces a
tank in the same layout; the tests use it, and it documents what the reader was validated against: there
is no vendor-written tank in this environment.
"""
from __future__ import annotations

import glob
import os
from types import SimpleNamespace
from typing import Dict, Optional

import numpy as np

TSQ_DTYPE = np.dtype([("size", "<i4"), ("type", "<i4"), ("code", "S4"), ("channel", "<u2"), ("sort_code", "<u2"),
                      ("timestamp", "<f8"), ("fp_loc", "<i8"), ("format", "<i4"), ("frequency", "<f4")])
assert TSQ_DTYPE.itemsize == 40

EVTYPE_STREAM = 0x00008101
EVTYPE_MASK = 0x0000FF0F          # stream headers may carry flag bits in the second byte
EVMARK_STARTBLOCK = 0x0001
EVMARK_STOPBLOCK = 0x0002
FORMATS = {0: np.dtype("<f4"), 1: np.dtype("<i4"), 2: np.dtype("<i2"), 3: np.dtype("<i1"),
           4: np.dtype("<f8"), 5: np.dtype("<i8")}


def _find(block_path: str, ext: str) -> str:
    hits = sorted(glob.glob(os.path.join(block_path, f"*.{ext}")))
    if not hits:
        raise FileNotFoundError(f"no .{ext} file in {block_path}")
    return hits[0]


def read_block(block_path: str, store: Optional[str] = None) -> SimpleNamespace:
    """``blk.streams.<NAME>.data`` (channels x samples), ``.fs``, ``.name``, ``.channels``, ``.start_time``;
    ``blk.info`` has the block's start / stop marks.  ``store``: read only this stream store."""
    heads = np.fromfile(_find(block_path, "tsq"), dtype=TSQ_DTYPE)
    if heads.size < 2:
        raise ValueError(f"{block_path}: the .tsq file holds no events")
    tev = np.memmap(_find(block_path, "tev"), dtype=np.uint8, mode="r")
    marks = {int.from_bytes(h["code"], "little"): float(h["timestamp"]) for h in heads
             if h["type"] != EVTYPE_STREAM and int.from_bytes(h["code"], "little") in (EVMARK_STARTBLOCK, EVMARK_STOPBLOCK)}
    is_stream = (heads["type"] & EVTYPE_MASK) == (EVTYPE_STREAM & EVTYPE_MASK)
    streams: Dict[str, SimpleNamespace] = {}
    for code in np.unique(heads["code"][is_stream]):
        name = code.decode("latin1").rstrip("\x00 ")
        if store is not None and name != store:
            continue
        h = heads[is_stream & (heads["code"] == code)]
        fmt = int(h["format"][0])
        if fmt not in FORMATS:
            raise ValueError(f"{block_path}: store {name} has unknown sample format {fmt}")
        dt = FORMATS[fmt]
        channels = np.unique(h["channel"])
        rows = []
        for ch in channels:
            hc = h[h["channel"] == ch]
            hc = hc[np.argsort(hc["timestamp"], kind="stable")]
            parts = []
            for rec in hc:
                nbytes = (int(rec["size"]) - 10) * 4
                lo = int(rec["fp_loc"])
                if nbytes < 0 or lo < 0 or lo + nbytes > tev.size:
                    raise ValueError(f"{block_path}: store {name} points outside the .tev file")
                parts.append(np.frombuffer(tev, dtype=dt, count=nbytes // dt.itemsize, offset=lo))
            rows.append(np.concatenate(parts) if parts else np.empty(0, dtype=dt))
        n = min(r.size for r in rows)                     # a block stopped mid-chunk: equal length rows
        data = np.stack([r[:n] for r in rows]).astype(dt.newbyteorder("="), copy=False)
        streams[name] = SimpleNamespace(name=name, data=data, fs=float(h["frequency"][0]),
                                        channels=[int(c) for c in channels], start_time=float(h["timestamp"].min()))
    return SimpleNamespace(streams=SimpleNamespace(**streams),
                           info=SimpleNamespace(start=marks.get(EVMARK_STARTBLOCK), stop=marks.get(EVMARK_STOPBLOCK),
                                                blockpath=block_path))


def write_block(block_path: str, streams: Dict[str, tuple], chunk: int = 256, tank: str = "tank",
                start_time: float = 1.6e9) -> None:
    """Write ``{store: (data (channels, samples), fs)}`` as a tank block in the layout above (test helper).
    Chunks of ``chunk`` samples, channels interleaved chunk by chunk like a live recording; the last chunk of
    a store is zero-padded to whole 32-bit words."""
    os.makedirs(block_path, exist_ok=True)
    block = os.path.basename(os.path.normpath(block_path))
    heads, payload, pos = [], [], 0

    def head(size, typ, code, ch, ts, loc, fmt, fs):
        r = np.zeros(1, dtype=TSQ_DTYPE)
        r["size"], r["type"], r["code"], r["channel"] = size, typ, code, ch
        r["timestamp"], r["fp_loc"], r["format"], r["frequency"] = ts, loc, fmt, fs
        return r

    heads.append(head(10, 0, (0).to_bytes(4, "little"), 0, 0.0, 0, 0, 0.0))
    heads.append(head(10, 0, EVMARK_STARTBLOCK.to_bytes(4, "little"), 0, start_time, 0, 0, 0.0))
    stop = start_time
    for name, (data, fs) in streams.items():
        data = np.atleast_2d(np.asarray(data))
        fmt = next(k for k, v in FORMATS.items() if v == data.dtype.newbyteorder("<"))
        for c0 in range(0, data.shape[1], chunk):
            for ch in range(data.shape[0]):
                raw = np.ascontiguousarray(data[ch, c0:c0 + chunk]).astype(data.dtype.newbyteorder("<")).tobytes()
                raw += b"\x00" * (-len(raw) % 4)
                ts = start_time + c0 / fs
                heads.append(head(10 + len(raw) // 4, EVTYPE_STREAM, name.encode("latin1").ljust(4, b"\x00")[:4], ch + 1, ts,
                                  pos, fmt, fs))
                payload.append(raw)
                pos += len(raw)
                stop = max(stop, ts + chunk / fs)
    heads.append(head(10, 0, EVMARK_STOPBLOCK.to_bytes(4, "little"), 0, stop, 0, 0, 0.0))
    np.concatenate(heads).tofile(os.path.join(block_path, f"{tank}_{block}.tsq"))
    with open(os.path.join(block_path, f"{tank}_{block}.tev"), "wb") as fh:
        for raw in payload:
            fh.write(raw)
