"""The C-ABI library loads and exports every symbol include/ecog_sm100.h declares.
No compute calls (there is no GPU in the CPU test tier)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ecog_sm100.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ecog_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound():
    from decode_tonal_langauge_b200 import _native as nat
    syms = declared_symbols()
    assert len(syms) >= 18
    raw = ctypes.CDLL(nat.LIB_PATH)
    for s in syms:
        assert hasattr(raw, s), f"{s} declared in the header but not exported"
        assert s in nat.PROTOTYPES, f"{s} has no ctypes prototype"
    assert set(nat.PROTOTYPES) == set(syms)
    assert nat.lib.ecog_abi_version() == nat.ABI_VERSION


def test_error_mapping_without_gpu():
    """Argument validation happens before any CUDA call, so it is testable on CPU."""
    from decode_tonal_langauge_b200 import _native as nat
    rc = nat.lib.ecog_car(None, None, 0, 10, 10, 10, None, 1.0, None)
    assert rc == nat.ECOG_E_VALUE
    with pytest.raises(ValueError):
        nat.check(rc)
    plan = nat.SosPlan(4, 1, 27, 1000, 1000)       # chunk not a multiple of 16
    rc = nat.lib.ecog_sosfilt(None, None, 2, 100, 100, 100, ctypes.byref(plan), ctypes.c_void_p(1), None, None,
                              None, 0, None)
    assert rc == nat.ECOG_E_VALUE


def test_round2_entry_points_validate_before_any_cuda_call():
    """ABI 7: the half-band pre-decimator and the float32 flag of the cascade pair reject what they are not built for."""
    import numpy as np
    from decode_tonal_langauge_b200 import _native as nat
    st = np.ones(30, dtype=np.float32)
    P = ctypes.c_void_p
    hp = st.ctypes.data_as(P)
    one = P(16)                                             # distinct dummy device pointers: nothing is launched
    rc = nat.lib.ecog_halfband2_decimate(one, P(32), 2, 4001, 4001, 1001, hp, 8, hp, 21, None)
    assert rc == nat.ECOG_E_VALUE                           # T not a multiple of 4
    rc = nat.lib.ecog_halfband2_decimate(one, P(32), 2, 4000, 4000, 1000, hp, 9, hp, 21, None)
    assert rc == nat.ECOG_E_UNSUPPORTED                     # more odd tap pairs than the kernel is built for
    with pytest.raises(NotImplementedError):
        nat.check(rc)
    rc = nat.lib.ecog_halfband2_decimate(one, one, 2, 4000, 4000, 1000, hp, 8, hp, 21, None)
    assert rc == nat.ECOG_E_VALUE                           # in place
    # the float32 band-pass half exists on the TMA path only
    sos = np.tile(np.array([1.0, 0.0, -1.0, 1.0, -1.5, 0.7]), (8, 1))
    zi = np.zeros((8, 2))
    plan = nat.SosPlan(8, 1, 27, 4096, 1024, nat.SOS_WARMUP, 256, 4 | nat.SOS_SPLIT_F32B, 512)
    ws = nat.lib.ecog_sos_workspace(ctypes.byref(plan), 2, 65536)
    rc = nat.lib.ecog_sosfilt(one, P(1 << 20), 2, 65536, 65536, 65536, ctypes.byref(plan), sos.ctypes.data_as(P),
                              zi.ctypes.data_as(P), None, P(1 << 24), ws, None)
    assert rc == nat.ECOG_E_VALUE


def test_ops_refuse_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from decode_tonal_langauge_b200 import ops
    with pytest.raises(RuntimeError):
        ops.require_cuda()
    with pytest.raises(TypeError):
        ops.car(torch.zeros(2, 8))          # CPU tensors are rejected, never silently processed
