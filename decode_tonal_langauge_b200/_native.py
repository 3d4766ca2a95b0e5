"""ctypes binding of libecog_sm100.so (see include/ecog_sm100.h).

There is no fallback: if the shared library is missing or an entry point is absent
this module raises at import time, and every op raises if CUDA is unavailable.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libecog_sm100.so")

ECOG_OK = 0
ECOG_E_VALUE = -1
ECOG_E_CUDA = -2
ECOG_E_WORKSPACE = -3
ECOG_E_UNSUPPORTED = -4
ABI_VERSION = 7
MAX_SECTIONS = 8
SOS_SCAN = 0
SOS_WARMUP = 1
SOS_WARMUP_TMA = 2
SOS_SPLIT_F32B = 0x100
HILBERT_N = 4096


class SosPlan(C.Structure):
    _fields_ = [("nsec", C.c_int32), ("zero_phase", C.c_int32), ("padlen", C.c_int32),
                ("chunk", C.c_int32), ("tail", C.c_int32), ("mode", C.c_int32), ("threads", C.c_int32),
                ("split", C.c_int32), ("tail_b", C.c_int32)]


class FftAxis(C.Structure):
    _fields_ = [("n", C.c_int32), ("nstage", C.c_int32), ("radix", C.c_int32 * 16)]


class ResamplePlan(C.Structure):
    _fields_ = [("T", C.c_int64), ("num", C.c_int64),
                ("fa", FftAxis), ("fb", FftAxis), ("ia", FftAxis), ("ib", FftAxis)]


class ResampleTables(C.Structure):
    _fields_ = [("tw_fa", C.c_void_p), ("tw_fb", C.c_void_p), ("tw_ia", C.c_void_p), ("tw_ib", C.c_void_p),
                ("tw_big_f_hi", C.c_void_p), ("tw_big_f_lo", C.c_void_p),
                ("tw_big_i_hi", C.c_void_p), ("tw_big_i_lo", C.c_void_p),
                ("big_f_split", C.c_int32), ("big_i_split", C.c_int32),
                ("tw_T", C.c_void_p), ("tw_num", C.c_void_p), ("bin_gain", C.c_void_p),
                ("tw_q_f", C.c_void_p), ("tw_q_i", C.c_void_p)]


class FftTables(C.Structure):
    _fields_ = [("tw_a", C.c_void_p), ("tw_b", C.c_void_p),
                ("tw_big_hi", C.c_void_p), ("tw_big_lo", C.c_void_p), ("tw_q", C.c_void_p)]


_P = C.c_void_p
_I64 = C.c_int64
_I32 = C.c_int32
_F64 = C.c_double
_SZ = C.c_size_t

# name -> (restype, argtypes); the CPU test suite checks every symbol declared in the header is here
PROTOTYPES = {
    "ecog_abi_version": (C.c_int, []),
    "ecog_last_error": (C.c_char_p, []),
    "ecog_launch_count": (_I64, []),
    "ecog_launch_log": (C.c_char_p, [C.c_int]),
    "ecog_car": (C.c_int, [_P, _P, _I64, _I64, _I64, _I64, _P, _F64, _P]),
    "ecog_car_colsum": (C.c_int, [_P, _I64, _I64, _I64, _P, _P, _P]),
    "ecog_car_apply": (C.c_int, [_P, _P, _I64, _I64, _I64, _I64, _P, _F64, _P]),
    "ecog_row_stats_workspace": (_SZ, [_I64, _I64]),
    "ecog_row_stats": (C.c_int, [_P, _I64, _I64, _I64, _I64, _I64, _P, _P, _P, _SZ, _P]),
    "ecog_zscore_apply": (C.c_int, [_P, _P, _I64, _I64, _I64, _I64, _P, _P, C.c_int, _P]),
    "ecog_sos_workspace": (_SZ, [C.POINTER(SosPlan), _I64, _I64]),
    "ecog_sosfilt": (C.c_int, [_P, _P, _I64, _I64, _I64, _I64, C.POINTER(SosPlan), _P, _P, _P, _P, _SZ, _P]),
    "ecog_copy2d": (C.c_int, [_P, _I64, _P, _I64, _I64, _I64, _P]),
    "ecog_hilbert_twiddle_floats": (_SZ, []),
    "ecog_hilbert_twiddles": (C.c_int, [_P]),
    "ecog_hilbert_env": (C.c_int, [_P, _P, _I64, _I64, _I64, _I64, _P, _I32, _I32, _P, _P, _I32, _I32, _P, _P, _F64, _P]),
    "ecog_hilbert_blocks": (_I64, [_I64, _I32]),
    "ecog_hilbert_env_range": (C.c_int, [_P, _P, _I64, _I64, _I64, _I64, _P, _I32, _I32, _P, _P, _I32, _I32, _P, _P, _F64,
                                         _I64, _I64, _P]),
    "ecog_resample_workspace": (_SZ, [C.POINTER(ResamplePlan), _I64]),
    "ecog_fft_resample": (C.c_int, [_P, _P, _I64, _I64, _I64, C.POINTER(ResamplePlan),
                                    C.POINTER(ResampleTables), _P, _SZ, _P]),
    "ecog_fft_c2c_workspace": (_SZ, [C.POINTER(FftAxis), C.POINTER(FftAxis), _I64]),
    "ecog_fft_c2c": (C.c_int, [_P, _P, _I64, _I64, _I64, C.POINTER(FftAxis), C.POINTER(FftAxis),
                               C.POINTER(FftTables), _I32, C.c_float, _P, _SZ, _P]),
    "ecog_cplx_modulate": (C.c_int, [_P, _I32, _I64, _I64, _P, _P, _I32, _I64, _I64, _I64, _P]),
    "ecog_cplx_abs_accumulate": (C.c_int, [_P, _I64, _P, _I64, _I64, _I64, _I32, C.c_float, _I32, _P]),
    "ecog_fir_decimate": (C.c_int, [_P, _P, _I64, _I64, _I64, _I64, _P, _I32, _I32, _I32, _P]),
    "ecog_halfband2_decimate": (C.c_int, [_P, _P, _I64, _I64, _I64, _I64, _P, _I32, _P, _I32, _P]),
    "ecog_fir_causal": (C.c_int, [_P, _P, _I64, _I64, _I64, _I64, _P, _I32, _P]),
    "ecog_rolling_workspace": (_SZ, [_I64, _I64]),
    "ecog_rolling_zscore": (C.c_int, [_P, _P, _I64, _I64, _I64, _I64, _I64, _P, C.c_int, _P, _SZ, _P]),
    "ecog_epoch_gather": (C.c_int, [_P, _P, _I64, _I64, _I64, _P, _P, _I64, _I64, _I32, _P]),
    "ecog_channel_select": (C.c_int, [_P, _P, _I64, _I64, _I64, _P, _P, _I64, _I32, _P]),
    "ecog_anova_workspace": (_SZ, [_I64, _I64, _I64, _I32]),
    "ecog_anova_f": (C.c_int, [_P, _I64, _P, _I64, _I64, _I64, _P, _P, _I32, _P, _P, _P, _SZ, _P]),
    "ecog_sig_runlength": (C.c_int, [_P, _I64, _I64, _F64, _P, _P]),
}


class NativeError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m decode_tonal_langauge_b200.build` "
            "(there is no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError here = stale build
        fn.restype = res
        fn.argtypes = args
    if lib.ecog_abi_version() != ABI_VERSION:
        raise ImportError(f"libecog_sm100 ABI {lib.ecog_abi_version()} != expected {ABI_VERSION}")
    return lib


lib = _load()


def check(rc: int) -> None:
    """Map a library return code to the reference's exception types."""
    if rc == ECOG_OK:
        return
    msg = (lib.ecog_last_error() or b"").decode("utf-8", "replace")
    if rc == ECOG_E_VALUE:
        raise ValueError(msg)
    if rc == ECOG_E_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise NativeError(f"libecog_sm100 error {rc}: {msg}")


def launch_log(reset: bool = True) -> list:
    """Names of the kernels this thread launched since the last reset (test / bench bookkeeping)."""
    raw = (lib.ecog_launch_log(1 if reset else 0) or b"").decode()
    return [k for k in raw.split(",") if k]


def launch_count() -> int:
    return int(lib.ecog_launch_count())
