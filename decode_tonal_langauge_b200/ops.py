"""Device operators: thin torch-tensor wrappers over the C ABI (include/ecog_sm100.h).

Every function takes float32 CUDA tensors of shape (C, T) (contiguous rows), enqueues
on the current torch stream and returns new CUDA tensors.  PyTorch only owns the memory
and the stream; all arithmetic happens in libecog_sm100.so.  No CPU path exists.
"""
from __future__ import annotations

import ctypes as C
import os
from collections import OrderedDict
from typing import Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native as nat
from . import design as D
from . import fftplan as FP

lib = nat.lib


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _hptr(a: Optional[np.ndarray]) -> C.c_void_p:
    return C.c_void_p(0 if a is None else a.ctypes.data)


def require_cuda() -> None:
    if not torch.cuda.is_available():
        raise RuntimeError("decode_tonal_langauge_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def as_signal(x: torch.Tensor) -> torch.Tensor:
    """(C, T) float32 CUDA tensor with contiguous rows (row stride may exceed T)."""
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise TypeError("expected a CUDA tensor")
    if x.dim() != 2:
        raise ValueError(f"expected a (channels, time) array, got shape {tuple(x.shape)}")
    if x.dtype != torch.float32:
        x = x.to(torch.float32)
    if x.stride(1) != 1 or (x.shape[0] > 1 and x.stride(0) < x.shape[1]):
        x = x.contiguous()
    return x


def _ld(x: torch.Tensor) -> int:
    return int(x.stride(0)) if x.shape[0] > 1 else int(x.shape[1])


_workspaces = {}


def workspace(nbytes: int, device, tag: str = "default") -> torch.Tensor:
    """Grow-only scratch buffer per (device, tag); owned by torch, handed to the library."""
    key = (str(device), tag)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        if buf is not None:
            torch.cuda.synchronize(device)      # rare: the old buffer may still be in use on another stream
        buf = None
        _workspaces.pop(key, None)
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def release_workspaces() -> None:
    """Drop every scratch buffer and every cached device table (they are rebuilt on demand)."""
    _workspaces.clear()
    _tables.clear()
    _pinned_tables.clear()
    _order_cache.clear()


# Device-side tables (twiddles, gains, FIR taps, per-length resample / chirp-z tables).  The
# per-length entries are O(T) each, and real recording lengths differ from block to block, so the
# cache is an LRU with a byte budget; small length-independent tables are pinned and never evicted.
TABLE_CACHE_BYTES = int(os.environ.get("ECOG_TABLE_CACHE_BYTES", str(1 << 30)))
_PINNED_TABLE_BYTES = 1 << 20
_tables: "OrderedDict" = OrderedDict()
_pinned_tables = {}
_table_bytes = 0


def _dev_table(key, make, device):
    k = (str(device), key)
    t = _pinned_tables.get(k)
    if t is not None:
        return t
    t = _tables.get(k)
    if t is not None:
        _tables.move_to_end(k)
        return t
    t = torch.from_numpy(np.ascontiguousarray(make())).to(device)
    nbytes = t.numel() * t.element_size()
    if nbytes <= _PINNED_TABLE_BYTES and key[0] in ("hilbert_tw", "hilbert_gain", "fir_taps"):
        _pinned_tables[k] = t
        return t
    _tables[k] = t
    total = sum(v.numel() * v.element_size() for v in _tables.values())
    while total > TABLE_CACHE_BYTES and len(_tables) > 1:
        _, old = _tables.popitem(last=False)         # least recently used; the caller still holds what it needs
        total -= old.numel() * old.element_size()
    return t


def table_cache_bytes() -> int:
    return sum(v.numel() * v.element_size() for v in _tables.values())


# ------------------------------------------------------------------------ K1
def car(x: torch.Tensor, exclude_channels: Sequence[int] = ()) -> torch.Tensor:
    x = as_signal(x)
    Cn, T = x.shape
    w, n_inc = _car_weights(Cn, exclude_channels, x.device)
    if Cn > CAR_FUSED_MAX_CHANNELS:
        # the [C x 32] strip of the fused kernel no longer fits in shared memory (e.g. 256 channels
        # after an 8-band frequency_filter): column sums, then subtract -- same arithmetic
        return car_apply(x, car_colsum(x, w), n_inc)
    y = torch.empty((Cn, T), dtype=torch.float32, device=x.device)
    nat.check(lib.ecog_car(_ptr(x), _ptr(y), Cn, T, _ld(x), _ld(y), _ptr(w), 1.0 / n_inc, _stream()))
    return y


# largest channel count whose [C x 32] float strip (+ partial sums) fits the fused kernel's 200 KB budget
CAR_FUSED_MAX_CHANNELS = (200 * 1024 // 4 - 32 * 32 - 32) // 32      # 1567 (csrc/car_zscore.cu, ecog_car)


def _car_weights(Cn, exclude_channels, device):
    if not isinstance(exclude_channels, (list, tuple)):
        raise ValueError("exclude_channels must be a list of integers.")
    if any(ch < 0 or ch >= Cn for ch in exclude_channels):
        raise ValueError("exclude_channels contains invalid channel indices.")
    if len(exclude_channels) == 0:
        return None, Cn
    w = np.ones(Cn, dtype=np.float32)
    w[list(exclude_channels)] = 0.0
    return torch.from_numpy(w).to(device), int(w.sum())


def car_colsum(x: torch.Tensor, weights: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Phase 1 of the channel-sharded CAR: per-timestep sum over this shard's rows."""
    x = as_signal(x)
    Cn, T = x.shape
    s = out if out is not None else torch.empty(T, dtype=torch.float32, device=x.device)
    if s.numel() != T or s.dtype != torch.float32 or not s.is_contiguous():
        raise ValueError("car_colsum: out must be a contiguous float32 vector of T elements")
    nat.check(lib.ecog_car_colsum(_ptr(x), Cn, T, _ld(x), _ptr(weights), _ptr(s), _stream()))
    return s


def car_apply(x: torch.Tensor, colsum: torch.Tensor, n_included: int) -> torch.Tensor:
    """Phase 2: subtract the (all-reduced) column sum divided by the global included count."""
    x = as_signal(x)
    Cn, T = x.shape
    y = torch.empty((Cn, T), dtype=torch.float32, device=x.device)
    nat.check(lib.ecog_car_apply(_ptr(x), _ptr(y), Cn, T, _ld(x), _ld(y), _ptr(colsum), 1.0 / n_included, _stream()))
    return y


# ------------------------------------------------------------------------ K2
def row_stats(x: torch.Tensor, t0: int = 0, t1: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    x = as_signal(x)
    Cn, T = x.shape
    t1 = T if t1 is None else t1
    mean = torch.empty(Cn, dtype=torch.float64, device=x.device)
    std = torch.empty(Cn, dtype=torch.float64, device=x.device)
    nbytes = lib.ecog_row_stats_workspace(Cn, T)
    ws = workspace(nbytes, x.device)
    nat.check(lib.ecog_row_stats(_ptr(x), Cn, T, _ld(x), int(t0), int(t1), _ptr(mean), _ptr(std),
                                 _ptr(ws), ws.numel(), _stream()))
    return mean, std


def zscore(x: torch.Tensor, t0: int = 0, t1: Optional[int] = None, nan_to_zero: bool = False) -> torch.Tensor:
    x = as_signal(x)
    mean, std = row_stats(x, t0, t1)
    y = torch.empty(tuple(x.shape), dtype=torch.float32, device=x.device)
    nat.check(lib.ecog_zscore_apply(_ptr(x), _ptr(y), x.shape[0], x.shape[1], _ld(x), _ld(y), _ptr(mean), _ptr(std),
                                    1 if nan_to_zero else 0, _stream()))
    return y


# ------------------------------------------------------------------------ K3
def _sos_threads_per_sm() -> int:
    return int(os.environ.get("ECOG_SOS_TPS", "512"))


def _pair_threads_per_sm() -> int:
    # 8 sections per thread carry twice the arithmetic per byte of a single cascade: two warps per scheduler
    # already fill the FP64 pipe, and half the threads means twice the chunk length, i.e. half the redundant
    # warm-up (measured at C2: 768 -> 12.3 ms, 512 -> 12.5, 384 -> 14.0, 256 -> 11.7)
    return int(os.environ.get("ECOG_PAIR_TPS", "256"))


def _tma_enabled() -> bool:
    """TMA-staged sweeps (csrc/sosfilt_tma.cu) whenever the recording allows them -- contiguous rows whose
    length has a usable divisor (tma_chunk); measured at C2 / one C4 shard: pair 11.9 -> 11.3 ms / 10.9 -> 8.9 ms,
    notch 8.8 -> 8.1 ms, band-pass 7.0 -> 6.4 ms.  ECOG_SOS_TMA=0 keeps the cp.async ring kernels."""
    return os.environ.get("ECOG_SOS_TMA", "1") != "0"


def _pair_f32_enabled() -> bool:
    return os.environ.get("ECOG_PAIR_F32", "1") != "0"


def tma_chunk(Cn: int, T: int, tail: int, max_per_sm: int = 2) -> Optional[int]:
    """Chunk length for the TMA sweeps (csrc/sosfilt_tma.cu): L divides T, L % 32 == 0, L >= tail.  The grid
    is ceil(Cn * T / L / 256) CTAs of 256 chunk-threads, two of which fit an SM; CTAs are dealt round robin,
    so the sweep lasts as long as the busiest SM's ceil(CTAs / 148) CTAs take, each doing L + tail samples per
    thread (a grid of 300 CTAs on 296 slots runs twice as long as one of 296).  One resident CTA leaves the
    FP64 pipe ~60 % busy, two ~75 % (ncu, profiles/r02_ncu_pair_tma.txt): the candidate with the least
    (CTAs per SM) x (L + tail) / utilisation wins; None if T has no such divisor."""
    best = None
    for n in range(2, T // max(tail, 32) + 1):
        if T % n:
            continue
        L = T // n
        if L % 32 or L < tail:
            continue
        ctas = -(-Cn * n // 256)
        per_sm = -(-ctas // D.NUM_SMS)
        if per_sm > max_per_sm:
            continue                       # more chunks than the resident CTAs per SM: only more warm-up
        cost = per_sm * (L + tail) / (0.62 if per_sm == 1 else 0.75)
        if best is None or cost < best[0]:
            best = (cost, L)
    return None if best is None else best[1]


PAIR_WS_CHUNKS = (256, 224, 192)      # chunks per CTA the warp-specialised pair may run with (csrc/sosfilt_pairws.cu)


def pair_ws_shape(Cn: int, T: int, tail: int) -> Optional[Tuple[int, int]]:
    """(chunks per CTA P, chunk length L) for the warp-specialised pair: one CTA per SM, so the sweep lasts
    ceil(CTAs / 148) x P x (L + tail) FP64-bound thread-samples per SM.  With P fixed at 256 the divisors of T leave
    125 CTAs for 148 SMs at C2; P = 224 makes it 143 CTAs of 7/8 the work each.  ECOG_PAIR_P pins P."""
    pin = os.environ.get("ECOG_PAIR_P")
    best = None
    for P in ((int(pin),) if pin else PAIR_WS_CHUNKS):
        for n in range(2, T // max(tail, 32) + 1):
            if T % n:
                continue
            L = T // n
            if L % 32 or L < tail:
                continue
            ctas = -(-Cn * n // P)
            waves = -(-ctas // D.NUM_SMS)
            cost = waves * P * (L + tail)
            if best is None or cost < best[0]:
                best = (cost, P, L)
    return None if best is None else (best[1], best[2])


def sosfilt(x: torch.Tensor, dsg: D.SosDesign, chunk: Optional[int] = None,
            out: Optional[torch.Tensor] = None, mode: Optional[str] = None, ws_tag: str = "sos") -> torch.Tensor:
    """Biquad cascade, zero-phase (filtfilt semantics) or causal.  ``mode``: "warm" (one kernel
    per sweep, start states from a zero-state warm-up), "scan" (exact carry scan) or None =
    warm-up whenever the cascade forgets fast enough for the chunk length."""
    x = as_signal(x)
    Cn, T = x.shape
    if dsg.zero_phase and T <= dsg.padlen:
        raise ValueError(f"The length of the input vector x must be greater than padlen, which is {dsg.padlen}.")
    y = out if out is not None else torch.empty((Cn, T), dtype=torch.float32, device=x.device)
    sos = np.ascontiguousarray(dsg.sos, dtype=np.float64)
    zi = None if dsg.zi is None else np.ascontiguousarray(dsg.zi, dtype=np.float64)
    Mh = None
    plan = None
    if (mode == "tma" or (mode is None and _tma_enabled())) and dsg.zero_phase and dsg.nsec == 4 \
            and _ld(x) == T and _ld(y) == T and x.data_ptr() % 16 == 0 and y.data_ptr() % 16 == 0:
        tail = D.warm_tail(dsg, T)
        tail32 = -(-tail // 32) * 32
        L = (int(chunk) if chunk is not None else tma_chunk(Cn, T, tail32)) if tail >= 0 else None
        if L is not None and T % L == 0 and L % 32 == 0 and tail32 <= L and T // L > 1:
            plan = nat.SosPlan(dsg.nsec, 1, dsg.padlen, L, tail32, nat.SOS_WARMUP_TMA, 256)
        elif mode == "tma":
            raise ValueError("the TMA sweeps need contiguous rows whose length has a divisor L >= tail with L % 32 == 0")
    if plan is None and mode in (None, "warm", "tma"):
        mode = None if mode == "tma" else mode
        tps = _sos_threads_per_sm()
        L = int(chunk) if chunk is not None else D.choose_warm_chunk(Cn, T, tps)
        if L % D.SUB:
            raise ValueError(f"chunk must be a multiple of {D.SUB}")
        n_chunks = -(-T // L)
        limit = 1 << 30 if mode == "warm" else int(D.WARM_MAX_OVERHEAD * L)
        tail = 0 if n_chunks == 1 else D.warm_tail(dsg, min(limit, T + D.SUB))
        if n_chunks > 1 and tail < 0 and mode == "warm":
            raise ValueError("the filter does not forget its state within the row; use mode='scan'")
        if tail >= 0:
            threads = 512 if Cn * n_chunks >= 2 * D.NUM_SMS * 512 else 256
            plan = nat.SosPlan(dsg.nsec, 1 if dsg.zero_phase else 0, dsg.padlen, L, tail, nat.SOS_WARMUP, threads)
    if plan is None:
        L = D.choose_chunk(Cn, T, chunk)
        n_chunks = -(-T // L)
        if n_chunks > 1:
            Mm, tail = D.chunk_ops(dsg, L)
            Mh = np.ascontiguousarray(Mm, dtype=np.float64)
        else:
            tail = L
        plan = nat.SosPlan(dsg.nsec, 1 if dsg.zero_phase else 0, dsg.padlen, L, tail, nat.SOS_SCAN, 512)
    nbytes = lib.ecog_sos_workspace(C.byref(plan), Cn, T)
    ws = workspace(nbytes, x.device, ws_tag)
    nat.check(lib.ecog_sosfilt(_ptr(x), _ptr(y), Cn, T, _ld(x), _ld(y), C.byref(plan), _hptr(sos), _hptr(zi),
                               _hptr(Mh), _ptr(ws), ws.numel(), _stream()))
    return y


_side_streams = {}


def _side_stream(device) -> "torch.cuda.Stream":
    key = str(device)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=device)
    return _side_streams[key]


def copy2d(dst: torch.Tensor, src: torch.Tensor) -> None:
    """dst[:, :] = src[:, :] for float32 row-major views (device-to-device, on the current stream)."""
    if dst.shape != src.shape or dst.dtype != torch.float32 or src.dtype != torch.float32:
        raise ValueError("copy2d needs two float32 tensors of the same shape")
    if dst.stride(1) != 1 or src.stride(1) != 1:
        raise ValueError("copy2d needs contiguous rows")
    nat.check(lib.ecog_copy2d(_ptr(src), _ld(src), _ptr(dst), _ld(dst), int(src.shape[0]), int(src.shape[1]), _stream()))


def pair_plan(Cn: int, T: int, ld_ok: bool, A: D.SosDesign, B: D.SosDesign):
    """Plan of the fused cascade pair for a (Cn, T) recording, or None when the pair cannot run
    (rows too short for the warm-up, non-unit numerators, misaligned rows)."""
    if not (A.zero_phase and B.zero_phase and A.nsec == 4 and B.nsec == 4 and ld_ok and T % 4 == 0):
        return None
    comb = D.pair_design(A, B)
    if comb is None:
        return None
    dsg, tail_b = comb
    L = D.choose_warm_chunk(Cn, T, _pair_threads_per_sm())
    n_chunks = -(-T // L)
    if n_chunks < 2:
        return None
    # forgetting time judged in the two filters' own state scaling (folding both gains onto the input
    # only rescales the first cascade's states; it does not make the pair remember longer)
    natural = D.SosDesign(np.ascontiguousarray(np.vstack([A.sos, B.sos])), None, dsg.padlen, True)
    tail = D.warm_tail(natural, min(int(D.WARM_MAX_OVERHEAD * L), T + D.SUB))
    if tail < 0 or 4 * tail + 64 > T:
        return None
    # one CTA per SM (the 8-section kernel holds 64 coefficient and 32 state registers per thread)
    tps = _pair_threads_per_sm()
    threads = 512 if tps >= 512 else (384 if tps == 384 else 256)
    plan = nat.SosPlan(8, 1, dsg.padlen, L, tail, nat.SOS_WARMUP, threads, 4, min(tail_b, tail))
    return dsg, plan, tail


# chunk length of the exact carry scan that recomputes the row ends of a cascade pair: the segments are
# short (2 x tail samples), so the scan wants many short chunks (2 C x 2 tail / 256 threads) rather than
# few long ones (12 launches of ~10 us instead of ~170 us each)
PAIR_EDGE_CHUNK = 256


def sosfilt_pair(x: torch.Tensor, A: D.SosDesign, B: D.SosDesign, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``filtfilt_B(filtfilt_A(x))`` (two consecutive zero-phase Butterworth steps of the reference chain,
    ref: frequency_filter.py:218-229 twice) in ONE forward and ONE backward sweep.

    Away from the row ends the four sweeps F_A, B_A, F_B, B_B are LTI and commute, so
    B_B F_B B_A F_A = (B_A B_B)(F_A F_B): an 8-section cascade per direction (ecog_sosfilt, split = 4).
    Within ``tail`` samples of a row end (``max|A^tail| < 1e-10`` for the pair) the padded start-ups do
    not commute; those samples are recomputed exactly, by the sequential four-sweep path on the
    first / last ``2 tail`` samples of every row (their inner halves are discarded), and overwrite the
    pair's result.  Falls back to the two sequential filtfilts when the pair does not apply."""
    x = as_signal(x)
    Cn, T = x.shape
    y = out if out is not None else torch.empty((Cn, T), dtype=torch.float32, device=x.device)
    ld_ok = _ld(x) % 4 == 0 and _ld(y) % 4 == 0 and x.data_ptr() % 16 == 0 and y.data_ptr() % 16 == 0
    pp = pair_plan(Cn, T, ld_ok, A, B)
    if pp is None:
        return sosfilt(sosfilt(x, A), B, out=out)
    dsg, plan, V = pp
    if _tma_enabled() and _ld(x) == T and _ld(y) == T:
        tail32 = -(-V // 32) * 32
        # band-pass second half in float32 delta form when its poles allow it (12 of 29 FP64 operations per
        # sample leave the FP64 pipe; <= ~1e-6 of the row maximum, ECOG_PAIR_F32=0 keeps everything in float64)
        f32b = _pair_f32_enabled() and D.bandpass_f32_ok(B)
        shape = pair_ws_shape(Cn, T, tail32) if f32b else None
        if shape is None:
            L = tma_chunk(Cn, T, tail32, max_per_sm=1)   # the 8-section kernel keeps its coefficients in registers: one CTA per SM
            shape = None if L is None else (256, L)
        if shape is not None and T // shape[1] > 1:
            split = 4 | (nat.SOS_SPLIT_F32B if f32b else 0)
            plan = nat.SosPlan(8, 1, dsg.padlen, shape[1], tail32, nat.SOS_WARMUP_TMA, shape[0], split,
                               -(-min(plan.tail_b, V) // 32) * 32)
    if T <= dsg.padlen:
        raise ValueError(f"The length of the input vector x must be greater than padlen, which is {dsg.padlen}.")
    # exact edges (they only read x): rows 0..C-1 = left segments, C..2C-1 = right segments.  Twelve small
    # launches on a side stream: they run beside the sweeps (whose grid leaves SMs free) instead of before them.
    E = 2 * V
    main = torch.cuda.current_stream()
    side = _side_stream(x.device)
    side.wait_stream(main)
    # the three edge buffers live in a persistent workspace (no allocator traffic across the two streams);
    # side.wait_stream(main) above also orders this call's writes behind the previous call's scatter
    nbuf = 2 * Cn * E
    ebuf = workspace(3 * nbuf * 4, x.device, "pair_edges")[:3 * nbuf * 4].view(torch.float32)
    xe, ze, ye = (ebuf[i * nbuf:(i + 1) * nbuf].view(2 * Cn, E) for i in range(3))
    with torch.cuda.stream(side):
        copy2d(xe[:Cn], x[:, :E])
        copy2d(xe[Cn:], x[:, T - E:])
        sosfilt(xe, A, chunk=PAIR_EDGE_CHUNK, mode="scan", ws_tag="sos_edge", out=ze)
        sosfilt(ze, B, chunk=PAIR_EDGE_CHUNK, mode="scan", ws_tag="sos_edge", out=ye)
    sos = np.ascontiguousarray(dsg.sos, dtype=np.float64)
    zi = np.ascontiguousarray(dsg.zi, dtype=np.float64)
    nbytes = lib.ecog_sos_workspace(C.byref(plan), Cn, T)
    ws = workspace(nbytes, x.device, "sos")
    nat.check(lib.ecog_sosfilt(_ptr(x), _ptr(y), Cn, T, _ld(x), _ld(y), C.byref(plan), _hptr(sos), _hptr(zi),
                               _hptr(None), _ptr(ws), ws.numel(), _stream()))
    main.wait_stream(side)
    copy2d(y[:, :V], ye[:Cn, :V])
    copy2d(y[:, T - V:], ye[Cn:, E - V:])
    return y


def butter(x: torch.Tensor, freqs, fs, order=4, causal=False, filter_type="bandpass",
           chunk: Optional[int] = None, out: Optional[torch.Tensor] = None, mode: Optional[str] = None) -> torch.Tensor:
    """ref: frequency_filter.py:187-229 (butter_filter keyword names kept)."""
    return sosfilt(x, D.butter_design(freqs, fs, order, causal, filter_type), chunk, out=out, mode=mode)


# ------------------------------------------------------------------------ K4
def _hilbert_twiddles(device) -> torch.Tensor:
    def make():
        tw = np.zeros(lib.ecog_hilbert_twiddle_floats(), dtype=np.float32)
        nat.check(lib.ecog_hilbert_twiddles(_hptr(tw)))
        return tw
    return _dev_table("hilbert_tw", make, device)


_hilbert_plans = {}


def hilbert(x: torch.Tensor, fs, freq_ranges, f0=0.018, octspace=1.0 / 7.0,
            filterbank_bias=np.log10(0.39), filterbank_slope=0.5, envelope=True,
            out: Optional[torch.Tensor] = None, car: Optional[Tuple[torch.Tensor, int]] = None,
            blocks: Optional[Tuple[int, int]] = None) -> torch.Tensor:
    """ref: frequency_filter.py:80-184 (hilbert_filter keyword names kept).

    ``car = (colsum, n_included)``: filter ``x - colsum / n_included`` instead of ``x`` -- a
    ``car_rereference`` step directly in front of the bank, folded into the kernel's load
    (ref: car_rereference.py:34-39; ``colsum`` from ``car_colsum``, all-reduced when channel-sharded).
    ``blocks = (b0, b1)``: only the overlap-save blocks [b0, b1) are computed (see ``hilbert_block_info``);
    needs ``out``."""
    x = as_signal(x)
    Cn, T = x.shape
    colsum, inv_count = (None, 0.0) if car is None else (car[0], 1.0 / car[1])
    if colsum is not None and (colsum.dtype != torch.float32 or colsum.numel() != T or not colsum.is_contiguous()):
        raise ValueError("car column sums must be a contiguous float32 vector of T elements")
    cfs, sds = D.gaussian_bank(freq_ranges, f0, octspace, filterbank_bias, filterbank_slope)
    if len(cfs) == 0:
        raise ValueError("the frequency ranges contain no filter-bank centre frequency")
    try:
        halo = FP.hilbert_halo(cfs, sds, float(fs), T)
    except NotImplementedError:
        if blocks is not None:
            raise ValueError("block ranges exist only on the block-wise path (hilbert_block_info returned None)")
        if colsum is not None:
            x = car_apply(x, colsum, car[1])
        return _hilbert_global(x, float(fs), cfs, sds, bool(envelope), out)      # low-frequency bands
    key = ("hilbert_gain", tuple(cfs.tolist()), tuple(sds.tolist()), float(fs), bool(envelope))
    plan = _hilbert_plans.get(key)
    if plan is None:
        g_h, sh, rw = FP.hilbert_gain(cfs, sds, float(fs), bool(envelope))
        plan = _hilbert_plans[key] = (g_h, sh, rw, FP.hilbert_nz(g_h))
    gain_h, shift, rows, nz = plan
    gain = _dev_table(key, lambda: gain_h, x.device)
    y = out if out is not None else torch.empty((Cn, T), dtype=torch.float32, device=x.device)
    if blocks is not None and out is None:
        raise ValueError("a block range writes into a caller-provided `out`")
    b0, b1 = (0, -1) if blocks is None else (int(blocks[0]), int(blocks[1]))
    nat.check(lib.ecog_hilbert_env_range(_ptr(x), _ptr(y), Cn, T, _ld(x), _ld(y), _ptr(gain), len(cfs), rows,
                                         _hptr(shift), _hptr(nz), halo, 1 if envelope else 0,
                                         _ptr(_hilbert_twiddles(x.device)), _ptr(colsum), float(inv_count), b0, b1,
                                         _stream()))
    return y


def hilbert_block_info(T: int, fs, freq_ranges, f0=0.018, octspace=1.0 / 7.0, filterbank_bias=np.log10(0.39),
                       filterbank_slope=0.5, **_):
    """(halo, useful samples per block U, number of blocks) of the block-wise bank for rows of T samples, or
    None when this bank runs on the whole-record path.  Block b writes samples [b U, (b+1) U) and reads
    [b U - halo, (b+1) U + halo) circularly."""
    cfs, sds = D.gaussian_bank(freq_ranges, f0, octspace, filterbank_bias, filterbank_slope)
    try:
        halo = FP.hilbert_halo(cfs, sds, float(fs), int(T))
    except NotImplementedError:
        return None
    return halo, nat.HILBERT_N - 2 * halo, int(lib.ecog_hilbert_blocks(int(T), halo))


HILBERT_GLOBAL_BLOCK_BYTES = 3 << 30


def _hilbert_global(x: torch.Tensor, fs: float, cfs, sds, envelope: bool, out: Optional[torch.Tensor]):
    """Whole-record Gaussian-Hilbert bank, exactly as the reference does it (ref:
    frequency_filter.py:154-184): FFT of the row, Gaussian x analytic mask per band, inverse FFT,
    |.| or real part, mean over bands.  Used when the bands' time kernels do not fit the 4096-sample
    blocks of the fused kernel (low-frequency bands).  Complex four-step FFTs of length T, or chirp-z
    convolutions when T is not a product of two 2-3-5 smooth factors (_hilbert_global_bluestein)."""
    Cn, T = x.shape
    try:
        bp = FP.big_plan(int(T), FP.MAX_AXIS_NARROW)
    except NotImplementedError:
        return _hilbert_global_bluestein(x, fs, cfs, sds, envelope, out)      # non-smooth row length
    dev = x.device
    nb = len(cfs)
    H = _global_band_gains(T, fs, cfs, sds)
    y = out if out is not None else torch.empty((Cn, T), dtype=torch.float32, device=dev)
    blk = int(max(1, min(Cn, HILBERT_GLOBAL_BLOCK_BYTES // (3 * T * 8))))
    for c0 in range(0, Cn, blk):
        c1 = min(Cn, c0 + blk)
        Z = torch.empty((c1 - c0, T, 2), dtype=torch.float32, device=dev)
        W = torch.empty_like(Z)
        nat.check(lib.ecog_cplx_modulate(_ptr(x[c0:c1]), 0, T, _ld(x), C.c_void_p(0), _ptr(Z), 1, T, T, c1 - c0,
                                         _stream()))
        fft_c2c(Z, bp)
        for b in range(nb):
            g = torch.from_numpy(FP._c2((H[b] / T).astype(np.complex128))).to(dev)     # 1/T of the inverse FFT
            _modulate(Z, T, g, W, T)
            fft_c2c(W, bp, inverse=True)
            nat.check(lib.ecog_cplx_abs_accumulate(_ptr(W), T, _ptr(y[c0:c1]), _ld(y), T, c1 - c0,
                                                   1 if envelope else 0, 1.0 / nb, 1 if b else 0, _stream()))
            del g
        del Z, W
    return y


def _global_band_gains(T: int, fs: float, cfs, sds) -> np.ndarray:
    """(nb, T) float64: Gaussian x analytic mask on the fftfreq grid, H[0] = 0 (ref: frequency_filter.py:155-175)."""
    freqs = np.fft.fftfreq(T, 1.0 / fs)
    h = np.zeros(T)
    if T % 2 == 0:
        h[0] = h[T // 2] = 1.0
        h[1:T // 2] = 2.0
    else:
        h[0] = 1.0
        h[1:(T + 1) // 2] = 2.0
    H = np.exp(-0.5 * ((freqs[None, :] - np.asarray(cfs)[:, None]) / np.asarray(sds)[:, None]) ** 2)
    H[:, 0] = 0.0
    return H * h[None, :]


def _hilbert_global_bluestein(x: torch.Tensor, fs: float, cfs, sds, envelope: bool, out: Optional[torch.Tensor]):
    """The whole-record bank for rows of ANY length (real TDT rates: 600 s at 3051.76 Hz = 1 831 054 = 2 x 915 527
    samples): both DFTs of frequency_filter.py:167,177 run as chirp-z convolutions of a smooth length
    M >= 2 T - 1 (fftplan.BluesteinPlan).  Per band: X g w -> FFT_M -> x FBi -> IFFT_M -> (x w) -> |.| or Re."""
    Cn, T = x.shape
    p = FP.bluestein_plan(int(T))
    dev = x.device
    nb = len(cfs)
    H = _global_band_gains(T, fs, cfs, sds)
    key = ("bluestein", T)
    pre = _dev_table(key + ("pre",), lambda: FP._c2(np.conj(p.w)), dev)
    FBf = _dev_table(key + ("FBf",), lambda: p.FBf, dev)
    FBi = _dev_table(key + ("FBi",), lambda: p.FBi, dev)
    post = _dev_table(key + ("post",), lambda: FP._c2(p.w), dev) if not envelope else None
    y = out if out is not None else torch.empty((Cn, T), dtype=torch.float32, device=dev)
    blk = int(max(1, min(Cn, HILBERT_GLOBAL_BLOCK_BYTES // (2 * p.M * 8 + T * 8))))
    for c0 in range(0, Cn, blk):
        c1 = min(Cn, c0 + blk)
        A = torch.empty((c1 - c0, p.M, 2), dtype=torch.float32, device=dev)
        _modulate(x[c0:c1], T, pre, A, p.M)                                   # x conj(w), zero padded
        fft_c2c(A, p.fft)
        _modulate(A, p.M, FBf, A, p.M)
        fft_c2c(A, p.fft, inverse=True)                                       # A[:T] conj(w) = X
        Z = torch.empty((c1 - c0, T, 2), dtype=torch.float32, device=dev)
        nat.check(lib.ecog_cplx_modulate(_ptr(A), 1, T, p.M, C.c_void_p(0), _ptr(Z), 1, T, T, c1 - c0, _stream()))
        for b in range(nb):
            # X[k] = A[k] conj(w[k]); the inverse wants X[k] g[k] w[k] = A[k] g[k] (|w| = 1): a REAL table
            g = torch.from_numpy(FP._c2((H[b] / T).astype(np.complex128))).to(dev)
            _modulate(Z, T, g, A, p.M)
            fft_c2c(A, p.fft)
            _modulate(A, p.M, FBi, A, p.M)
            fft_c2c(A, p.fft, inverse=True)
            if not envelope:
                _modulate(A, T, post, A, T)                                   # the unit phasor w[n] matters for Re
            nat.check(lib.ecog_cplx_abs_accumulate(_ptr(A), p.M, _ptr(y[c0:c1]), _ld(y), T, c1 - c0,
                                                   1 if envelope else 0, 1.0 / nb, 1 if b else 0, _stream()))
            del g
        del A, Z
    return y


# ------------------------------------------------------------------ K6 / K7
def fir_bank_taps(fs, order: int, center_frequencies) -> np.ndarray:
    """Averaged taps of the reference's FIR bank (ref: frequency_filter.py:260-274).  The
    reference divides the band edges by the Nyquist frequency AND passes ``fs`` to firwin
    (:265-268); reproduced as is.  float64."""
    from scipy import signal as sp_signal
    cfs = list(center_frequencies)
    if not cfs:
        raise ZeroDivisionError("division by zero")      # the reference divides by len(center_frequencies)
    nyq = 0.5 * float(fs)
    h = np.zeros(int(order) + 1, dtype=np.float64)
    for fc in cfs:
        h += sp_signal.firwin(int(order) + 1, [fc * 0.9 / nyq, fc * 1.1 / nyq], pass_zero=False, fs=fs)
    return h / len(cfs)


def fir_causal(x: torch.Tensor, h: np.ndarray, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y[c, t] = sum_j h[j] x[c, t - j] with zero initial state (scipy lfilter(h, 1, x))."""
    x = as_signal(x)
    Cn, T = x.shape
    h = np.asarray(h, dtype=np.float64)
    n4 = -(-(len(h) - 1) // 4) + 1
    off = 4 * n4 - 4
    g = np.zeros(4 * n4, dtype=np.float32)
    idx = off - np.arange(4 * n4)
    ok = (idx >= 0) & (idx < len(h))
    g[ok] = h[idx[ok]].astype(np.float32)
    gd = _dev_table(("fir_taps", g.tobytes()), lambda: g, x.device)
    y = out if out is not None else torch.empty((Cn, T), dtype=torch.float32, device=x.device)
    nat.check(lib.ecog_fir_causal(_ptr(x), _ptr(y), Cn, T, _ld(x), _ld(y), _ptr(gd), n4, _stream()))
    return y


def fir_bank(x: torch.Tensor, fs, order: int, center_frequencies, out: Optional[torch.Tensor] = None):
    """ref: frequency_filter.py:232-274 (fir_bandpass_filter keyword names kept)."""
    return fir_causal(x, fir_bank_taps(fs, order, center_frequencies), out=out)


def rolling_zscore(x: torch.Tensor, window: int, nan_to_zero: bool = False) -> torch.Tensor:
    """ref: rolling_zscore.py:28-49: trailing window of ``window`` samples, min_periods=1, ddof=1."""
    x = as_signal(x)
    Cn, T = x.shape
    mean, _ = row_stats(x)
    nbytes = lib.ecog_rolling_workspace(Cn, T)
    ws = workspace(nbytes, x.device, "rolling")
    y = torch.empty((Cn, T), dtype=torch.float32, device=x.device)
    nat.check(lib.ecog_rolling_zscore(_ptr(x), _ptr(y), Cn, T, _ld(x), _ld(y), int(window), _ptr(mean),
                                      1 if nan_to_zero else 0, _ptr(ws), ws.numel(), _stream()))
    return y


# ------------------------------------------------------------------------ K5
def _axis(ap: FP.AxisPlan) -> nat.FftAxis:
    a = nat.FftAxis()
    a.n = ap.n
    a.nstage = len(ap.radices)
    for i, r in enumerate(ap.radices):
        a.radix[i] = r
    return a


def fir_decimate(x: torch.Tensor, taps: np.ndarray, offset: int, D: int) -> torch.Tensor:
    """y[c, m] = sum_j taps[j] x[c, (m D + j - offset) mod T] (circular), float32."""
    x = as_signal(x)
    Cn, T = x.shape
    taps = np.ascontiguousarray(taps, dtype=np.float32)
    y = torch.empty((Cn, T // int(D)), dtype=torch.float32, device=x.device)
    nat.check(lib.ecog_fir_decimate(_ptr(x), _ptr(y), Cn, T, _ld(x), _ld(y), _hptr(taps), int(taps.shape[0]),
                                    int(offset), int(D), _stream()))
    return y


def halfband2_decimate(x: torch.Tensor, stage1: np.ndarray, stage2: np.ndarray) -> torch.Tensor:
    """Decimation by 4 as two circular half-band stages (ecog_halfband2_decimate); ``stage`` = centre tap, odd taps."""
    x = as_signal(x)
    Cn, T = x.shape
    s1 = np.ascontiguousarray(stage1, dtype=np.float32)
    s2 = np.ascontiguousarray(stage2, dtype=np.float32)
    y = torch.empty((Cn, T // 4), dtype=torch.float32, device=x.device)
    nat.check(lib.ecog_halfband2_decimate(_ptr(x), _ptr(y), Cn, T, _ld(x), _ld(y), _hptr(s1), int(s1.size) - 1,
                                          _hptr(s2), int(s2.size) - 1, _stream()))
    return y


def _fft_resample(x: torch.Tensor, num: int, bin_gain: Optional[np.ndarray], gain_key=None) -> torch.Tensor:
    Cn, T = x.shape
    rp = FP.resample_plan(int(T), num)
    dev = x.device
    key = ("resample", T, num)
    t = lambda name, arr: _dev_table(key + (name,), lambda: arr, dev)
    tables = nat.ResampleTables()
    keep = []
    for name, arr in (("tw_fa", rp.fwd.a.tw), ("tw_fb", rp.fwd.b.tw),
                      ("tw_ia", rp.inv.a.tw), ("tw_ib", rp.inv.b.tw), ("tw_big_f_hi", rp.fwd.tw_hi),
                      ("tw_big_f_lo", rp.fwd.tw_lo), ("tw_big_i_hi", rp.inv.tw_hi), ("tw_big_i_lo", rp.inv.tw_lo),
                      ("tw_T", rp.tw_T), ("tw_num", rp.tw_num), ("tw_q_f", rp.fwd.tw_q), ("tw_q_i", rp.inv.tw_q)):
        dt = t(name, arr)
        keep.append(dt)
        setattr(tables, name, dt.data_ptr())
    if bin_gain is not None:
        g = _dev_table(("resample_gain", gain_key, num), lambda: bin_gain, dev)
        keep.append(g)
        tables.bin_gain = g.data_ptr()
    tables.big_f_split = FP.BIG_SPLIT
    tables.big_i_split = FP.BIG_SPLIT
    plan = nat.ResamplePlan()
    plan.T, plan.num = T, num
    plan.fa, plan.fb, plan.ia, plan.ib = _axis(rp.fwd.a), _axis(rp.fwd.b), _axis(rp.inv.a), _axis(rp.inv.b)
    if _ld(x) % 2 or x.data_ptr() % 8:
        x = x.contiguous()
    nbytes = lib.ecog_resample_workspace(C.byref(plan), Cn)
    ws = workspace(nbytes, dev, "fft")
    y = torch.empty((Cn, num), dtype=torch.float32, device=dev)
    nat.check(lib.ecog_fft_resample(_ptr(x), _ptr(y), Cn, _ld(x), _ld(y), C.byref(plan), C.byref(tables),
                                    _ptr(ws), ws.numel(), _stream()))
    return y


def _fft_tables(bp: FP.BigPlan, key, dev):
    t = lambda name, arr: _dev_table(key + (name,), lambda: arr, dev)
    tb = nat.FftTables()
    keep = [t("tw_a", bp.a.tw), t("tw_b", bp.b.tw),
            t("tw_hi", bp.tw_hi), t("tw_lo", bp.tw_lo), t("tw_q", bp.tw_q)]
    tb.tw_a, tb.tw_b, tb.tw_big_hi, tb.tw_big_lo, tb.tw_q = [k.data_ptr() for k in keep]
    return tb, keep


def fft_c2c(z: torch.Tensor, bp: FP.BigPlan, inverse: bool = False, scale: float = 1.0,
            out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Complex FFT of every row of ``z`` ((C, N, 2) float32, N = bp.N), natural order."""
    Cn, N, _ = z.shape
    if N != bp.N:
        raise ValueError(f"row length {N} does not match the plan ({bp.N})")
    out = z if out is None else out
    tb, keep = _fft_tables(bp, ("c2c", bp.N), z.device)
    fa, fb = _axis(bp.a), _axis(bp.b)
    ws = workspace(lib.ecog_fft_c2c_workspace(C.byref(fa), C.byref(fb), Cn), z.device, "fft")
    nat.check(lib.ecog_fft_c2c(_ptr(z), _ptr(out), Cn, int(z.stride(0)) // 2, int(out.stride(0)) // 2, C.byref(fa),
                               C.byref(fb), C.byref(tb), 1 if inverse else 0, float(scale), _ptr(ws), ws.numel(),
                               _stream()))
    return out


def _modulate(src: torch.Tensor, in_len: int, table: torch.Tensor, dst: torch.Tensor, out_len: int) -> None:
    """dst[c, i] = (i < in_len ? src[c, i] : 0) * table[i]; complex tensors are (C, n, 2) float32."""
    in_c, out_c = src.dim() == 3, dst.dim() == 3
    ld_in = int(src.stride(0)) // (2 if in_c else 1)
    ld_out = int(dst.stride(0)) // (2 if out_c else 1)
    nat.check(lib.ecog_cplx_modulate(_ptr(src), 1 if in_c else 0, int(in_len), ld_in, _ptr(table), _ptr(dst),
                                     1 if out_c else 0, int(out_len), ld_out, int(src.shape[0]), _stream()))


CZT_BLOCK_BYTES = 2 << 30        # channel block size of the chirp-z path (two complex buffers)


def _czt_resample(x: torch.Tensor, num: int, bin_gain: Optional[np.ndarray] = None, gain_key=None) -> torch.Tensor:
    """Any-length scipy.signal.resample via Bluestein (plan and tables: fftplan.czt_plan)."""
    Cn, T = x.shape
    p = FP.czt_plan(int(T), int(num))
    dev = x.device
    key = ("czt", T, num)
    mid_h = p.mid if bin_gain is None else p.mid * np.asarray(bin_gain[:p.K], dtype=np.float64)
    tab = {name: _dev_table(key + (name,), (lambda a=arr: a), dev)
           for name, arr in (("pre", p.pre), ("FB1", p.FB1), ("FB2", p.FB2), ("post", p.post))}
    tab["mid"] = _dev_table(key + ("mid", gain_key), lambda: FP._c2(mid_h), dev)
    y = torch.empty((Cn, num), dtype=torch.float32, device=dev)
    M = max(p.M1, p.M2)
    blk = int(max(1, min(Cn, CZT_BLOCK_BYTES // (2 * M * 8))))
    for c0 in range(0, Cn, blk):
        c1 = min(Cn, c0 + blk)
        a = torch.empty((c1 - c0, p.M1, 2), dtype=torch.float32, device=dev)
        _modulate(x[c0:c1], T, tab["pre"], a, p.M1)                      # x conj(w), zero padded
        fft_c2c(a, p.fft1)
        _modulate(a, p.M1, tab["FB1"], a, p.M1)                          # times FFT of the chirp kernel
        fft_c2c(a, p.fft1, inverse=True)
        b = torch.empty((c1 - c0, p.M2, 2), dtype=torch.float32, device=dev)
        _modulate(a, p.K, tab["mid"], b, p.M2)                           # resample rules, inverse chirp, padded
        del a
        fft_c2c(b, p.fft2)
        _modulate(b, p.M2, tab["FB2"], b, p.M2)
        fft_c2c(b, p.fft2, inverse=True)
        _modulate(b, num, tab["post"], y[c0:c1], num)                    # Re(. v[m])
        del b
    return y


def fft_resample(x: torch.Tensor, num: int, two_stage: Optional[bool] = None) -> torch.Tensor:
    """scipy.signal.resample(x, num, axis=1) for real rows (ref: downsample.py:21-27).

    Large down-sampling ratios run in two stages: a circular FIR low-pass + decimate by D
    (ecog_fir_decimate), then the brick wall on the T/D-sample rows with the FIR's pass-band
    response divided out bin by bin.  ``two_stage=False`` forces the single whole-row FFT."""
    x = as_signal(x)
    Cn, T = x.shape
    num = int(num)
    if num < 2 or T < 2:
        raise ValueError("resample needs at least two samples in and out")

    def smooth(n_in: int) -> bool:
        try:
            FP.resample_plan(n_in, num)
            return True
        except NotImplementedError:
            return False

    pre = FP.predecimation(int(T), num) if two_stage in (None, True) else None
    if pre is not None:
        x1 = halfband2_decimate(x, *pre.halfband) if pre.halfband is not None else fir_decimate(x, pre.taps, pre.offset, pre.D)
        if smooth(int(T) // pre.D):
            return _fft_resample(x1, num, pre.bin_gain, gain_key=(int(T), pre.D))
        return _czt_resample(x1, num, pre.bin_gain, gain_key=(int(T), pre.D))     # non-smooth T/D: Bluestein
    if smooth(int(T)):
        return _fft_resample(x, num, None)
    return _czt_resample(x, num)


# ------------------------------------------------------------------------ K8
def epoch_gather(src: torch.Tensor, starts: np.ndarray, length: int) -> torch.Tensor:
    """out[n, c, :] = src[c, starts[n] : starts[n] + length]; bit copy for 1-, 2-, 4- and 8-byte dtypes
    (the reference slices whatever dtype the .npz holds, e.g. int16 audio)."""
    if not src.is_cuda or src.dim() != 2:
        raise TypeError("expected a (channels, time) CUDA tensor")
    if src.element_size() not in (1, 2, 4, 8):
        raise TypeError(f"epoch_gather supports 1-, 2-, 4- and 8-byte element types, got {src.dtype}")
    if src.stride(1) != 1:
        src = src.contiguous()
    Cn, T = src.shape
    starts = np.ascontiguousarray(starts, dtype=np.int64)
    N = int(starts.shape[0])
    out = torch.empty((N, Cn, int(length)), dtype=src.dtype, device=src.device)
    d_start = torch.from_numpy(starts).to(src.device) if N else None
    nat.check(lib.ecog_epoch_gather(_ptr(src), _ptr(out), Cn, T, _ld(src), _ptr(d_start), _hptr(starts), N,
                                    int(length), src.element_size(), _stream()))
    return out


def channel_select(epochs: torch.Tensor, channels) -> torch.Tensor:
    """out[n, j, :] = epochs[n, channels[j], :] (bit copy, 1-, 2-, 4- and 8-byte dtypes)."""
    if not epochs.is_cuda or epochs.dim() != 3:
        raise TypeError("expected an (events, channels, time) CUDA tensor")
    if epochs.element_size() not in (1, 2, 4, 8):
        raise TypeError(f"channel_select supports 1-, 2-, 4- and 8-byte element types, got {epochs.dtype}")
    epochs = epochs.contiguous()
    N, Cn, L = epochs.shape
    ch = np.ascontiguousarray(channels, dtype=np.int64)
    if ch.ndim != 1:
        raise IndexError("channel indices must be one-dimensional")
    ch = np.where(ch < 0, ch + Cn, ch).astype(np.int32)              # numpy-style negative indices
    K = int(ch.shape[0])
    out = torch.empty((N, K, L), dtype=epochs.dtype, device=epochs.device)
    d_ch = torch.from_numpy(ch).to(epochs.device) if K else None
    try:
        nat.check(lib.ecog_channel_select(_ptr(epochs), _ptr(out), N, Cn, L, _ptr(d_ch), _hptr(ch), K,
                                          epochs.element_size(), _stream()))
    except ValueError as e:
        raise IndexError(str(e))                                      # numpy raises IndexError here
    return out


# -------------------------------------------------------------------- K9 / K10
def anova_f(epochs: torch.Tensor, groups: np.ndarray, extra: Optional[torch.Tensor] = None):
    """One-way ANOVA over events for every (channel, timepoint).

    epochs (Na, C, L) [+ extra (Nb, C, L), concatenated after it]; groups: int array of
    length Na + Nb with values 0..G-1.  Returns (F, p) as (C, L) float64 CUDA tensors.
    The kernel reads float32 epochs and accumulates in float64 about the first event's value
    (float64 accumulators of float32 samples match scipy's f_oneway to 5e-13); float64
    epochs are narrowed on the device first -- their 29 extra mantissa bits are below the statistic's
    own conditioning, and the reference's epochs are float64 only because filtfilt promotes."""
    def prep(e):
        if e.dtype != torch.float32:
            e = e.to(torch.float32)
        return e.contiguous()
    epochs = prep(epochs)
    Na, Cn, L = epochs.shape
    Nb = 0
    if extra is not None:
        extra = prep(extra)
        Nb = int(extra.shape[0])
        if tuple(extra.shape[1:]) != (Cn, L):
            raise ValueError(f"Shape mismatch between recordings: {tuple(epochs.shape[1:])} vs {tuple(extra.shape[1:])}.")
    groups = np.ascontiguousarray(groups, dtype=np.int32)
    if groups.shape[0] != Na + Nb:
        raise ValueError("one group label per event is required")
    G, counts, d_groups = _group_order(groups, epochs.device)
    F = torch.empty((Cn, L), dtype=torch.float64, device=epochs.device)
    P = torch.empty((Cn, L), dtype=torch.float64, device=epochs.device)
    ws = workspace(lib.ecog_anova_workspace(Cn, L, Na + Nb, max(G, 2)), epochs.device, "anova")
    nat.check(lib.ecog_anova_f(_ptr(epochs), Na, _ptr(extra), Nb, Cn, L, _ptr(d_groups), _hptr(counts), G,
                               _ptr(F), _ptr(P), _ptr(ws), ws.numel(), _stream()))
    return F, P


_order_cache: "OrderedDict" = OrderedDict()


def _group_order(groups: np.ndarray, device):
    """(G, per-group counts, device tensor of the event numbers sorted by group).  The stable argsort and
    its upload depend on the label vector only: the last few are kept, so that the three selections of one
    sample set (tone, syllable, active) and repeated calls do not pay them on the critical path again."""
    key = (str(device), groups.shape[0], hash(groups.tobytes()))
    hit = _order_cache.get(key)
    if hit is not None and np.array_equal(hit[3], groups):
        _order_cache.move_to_end(key)
        return hit[0], hit[1], hit[2]
    G = int(groups.max()) + 1 if groups.size else 0
    counts = np.bincount(groups, minlength=max(G, 1)).astype(np.int64)
    order = np.argsort(groups, kind="stable").astype(np.int32)       # events sorted by group
    staged = torch.from_numpy(order).pin_memory()
    d_order = staged.to(device, non_blocking=True)
    _order_cache[key] = (G, counts, d_order, groups.copy(), staged)
    while len(_order_cache) > 8:
        _order_cache.popitem(last=False)
    return G, counts, d_order


def to_host_many(*tensors: torch.Tensor):
    """Several small device results -> numpy with ONE synchronisation (pinned staging, async copies)."""
    outs = []
    for t in tensors:
        t = t.contiguous()
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t, non_blocking=True)
        outs.append(h)
    torch.cuda.current_stream().synchronize()
    return [h.numpy() for h in outs]


def sig_runlength(p: torch.Tensor, threshold: float) -> torch.Tensor:
    """Longest run of consecutive p < threshold per channel (int32, (C,))."""
    p = p.contiguous()
    Cn, L = p.shape
    out = torch.empty(Cn, dtype=torch.int32, device=p.device)
    nat.check(lib.ecog_sig_runlength(_ptr(p), Cn, L, float(threshold), _ptr(out), _stream()))
    return out
