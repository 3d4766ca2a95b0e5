"""Round-2 parity tests on the B200: the shapes the bench runs (full-length rows, default plans),
the fused cascade pair and the CAR-in-Hilbert fold against the step-by-step path and the oracle,
strided views through every operator, and the widened operator limits (2048-row CAR, int16
gathers, more than 16 ANOVA groups)."""
from argparse import Namespace

import numpy as np
import pytest

from conftest import max_rel

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
TOL = 1e-5          # max_t|y - ref| / max_t|ref| per channel


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def _launched(fn):
    """(result, names of the library kernels `fn` launched) -- the library's own launch log."""
    from decode_tonal_langauge_b200 import _native as nat
    nat.launch_log(reset=True)
    out = fn()
    torch.cuda.synchronize()
    return out, set(nat.launch_log(reset=True))


def oracle_rows_with_car(x_dev, rows, fs, steps_without_car):
    """The oracle's chain for `rows` of a device-resident recording whose chain contains ONE
    car_rereference: CAR is linear and identical at every sample, and every step in front of it is a
    per-row linear filter, so  chain(x)[c] = chain_without_car(x[c] - mean_k x[k]).  The column mean
    is taken in float64; the oracle then sees float64 rows (the reference computes in float64 from
    its first filtfilt on).  Returns (float64 oracle result, long-double-notch result)."""
    from oracle import chains as CH
    from oracle import steps as OS
    from decode_tonal_langauge_b200 import design as D
    m = x_dev.mean(dim=0, dtype=torch.float64).cpu().numpy()
    xin = x_dev[rows].to(torch.float64).cpu().numpy() - m[None]
    ref, f_ref = CH.run_chain(xin, fs, steps_without_car)
    d = D.butter_design([58, 62], fs, 4, False, "bandstop")
    notch_ld = np.asarray(OS.filtfilt_pad(d.b, d.a, xin, dtype=np.longdouble, reference_edges=True), dtype=np.float64)
    truth, _ = CH.run_chain(notch_ld, fs, steps_without_car[1:])
    return ref, truth, f_ref


@pytest.mark.parametrize("C,T,fs", [(256, 7_200_000, 2000), (128, 10_800_000, 3000)],
                         ids=["C2-2kHz-60min", "C4shard-3kHz-60min"])
def test_full6_full_length_default_plans_vs_oracle(C, T, fs):
    """FULL6 on the BASELINE shapes themselves (configs[1]: 256 x 7.2 M @ 2 kHz; one 8-way shard of
    configs[3]: 128 x 10.8 M @ 3 kHz with the 58-62 Hz notch at pole radius 0.9985), default plan
    selection -- the cascade-pair warm-up kernel, the CAR fold, the two-stage resampler -- against
    the oracle on three rows, with the long-double rule for the notch."""
    from decode_tonal_langauge_b200 import synth
    from decode_tonal_langauge_b200.chains import FULL6_STEPS
    from decode_tonal_langauge_b200.preprocessor import preprocess_signal
    x = synth.device_session(C, T, fs, seed=11)
    (y, f), kernels = _launched(lambda: preprocess_signal(x, FULL6_STEPS, Namespace(signal_freq=fs)))
    names = " ".join(kernels)
    assert "sos_pair_ws_fwd" in kernels and "sos_pair_ws_bwd" in kernels, names   # the fused pair (TMA tiles, notch and band-pass threads), not a fallback
    assert "halfband2_decimate" in kernels, names            # two-stage resampler (two half-band stages in front of the FFT)
    assert "hilbert_env8" in kernels and "car_colsum" in kernels and "car_fused" not in kernels, names
    rows = [0, C // 2 + 1, C - 1]
    no_car = [FULL6_STEPS[0]] + FULL6_STEPS[2:]
    ref, truth, f_ref = oracle_rows_with_car(x, rows, fs, no_car)
    got = y[rows].cpu().numpy()
    assert f == f_ref == 400 and got.shape == ref.shape
    err_gpu, err_ref = max_rel(got, truth), max_rel(ref, truth)
    print(f"FULL6 {C}x{T}@{fs}: gpu vs long-double truth {err_gpu:.2e}, float64 reference vs truth {err_ref:.2e}, "
          f"gpu vs float64 reference {max_rel(got, ref):.2e}")
    assert err_gpu <= max(TOL, err_ref), (err_gpu, err_ref)
    # and the step-by-step path (no fusion) gives the same rows
    y0, _ = preprocess_signal(x, FULL6_STEPS, Namespace(signal_freq=fs), fuse=False)
    assert max_rel(y0[rows].cpu().numpy(), got) < 5e-6


def test_full6_three_channels_full_length_direct():
    """The verdict's direct form: FULL6 on 3 channels x 7.2 M against oracle.run_chain on the very
    same array (CAR over the three rows), warm-up IIR forced."""
    from oracle import chains as CH
    from oracle import steps as OS
    from decode_tonal_langauge_b200 import design as D
    from decode_tonal_langauge_b200 import ops, synth
    from decode_tonal_langauge_b200.chains import FULL6_STEPS
    T, fs = 7_200_000, 2000
    x = synth.session_channels(3, range(3), T, fs, 8)
    xd = torch.from_numpy(x).cuda()
    a = ops.butter(xd, [58, 62], fs, filter_type="bandstop", mode="warm", chunk=24336)
    b = ops.car(a)
    c = ops.butter(b, [70, 150], fs, filter_type="bandpass", mode="warm", chunk=24336)
    e = ops.hilbert(c, fs, [70.0, 150.0])
    r = ops.fft_resample(e, T // 5)
    z = ops.zscore(r).cpu().numpy()
    ref, _ = CH.run_chain(x, fs, FULL6_STEPS)
    d = D.butter_design([58, 62], fs, 4, False, "bandstop")
    notch_ld = np.asarray(OS.filtfilt_pad(d.b, d.a, x, dtype=np.longdouble, reference_edges=True), dtype=np.float64)
    truth, _ = CH.run_chain(notch_ld, fs, FULL6_STEPS[1:])
    err_gpu, err_ref = max_rel(z, truth), max_rel(ref, truth)
    print(f"3 ch x 7.2 M: gpu vs truth {err_gpu:.2e}, reference vs truth {err_ref:.2e}")
    assert err_gpu <= max(TOL, err_ref), (err_gpu, err_ref)
    # per step against the float64 oracle where the step is well conditioned
    assert max_rel(b.cpu().numpy(), OS.car_rereference(a.cpu().numpy().astype(np.float64))) < 2e-6
    bp_ref = OS.butter_filter(b.cpu().numpy(), [70, 150], fs, filter_type="bandpass")
    assert max_rel(c.cpu().numpy(), bp_ref) < TOL


@pytest.mark.parametrize("fs,T", [(2000, 4_400_000), (3000, 7_000_000)])
def test_cascade_pair_equals_sequential_and_oracle(fs, T, monkeypatch):
    """ops.sosfilt_pair = band-pass(filtfilt) o notch(filtfilt): against the two sequential device
    filtfilts on every row (interior AND the recomputed row ends) and against scipy on two rows."""
    from oracle import steps as OS
    from decode_tonal_langauge_b200 import design as D
    from decode_tonal_langauge_b200 import ops, synth
    C = 256
    x = synth.device_session(C, T, fs, seed=5)
    A = D.butter_design([58, 62], fs, 4, False, "bandstop")
    B = D.butter_design([70, 150], fs, 4, False, "bandpass")
    assert ops.pair_plan(C, T, True, A, B) is not None
    monkeypatch.setenv("ECOG_SOS_TMA", "0")                         # the cp.async ring kernels
    y, kernels = _launched(lambda: ops.sosfilt_pair(x, A, B))
    assert "sos_warm_pair_fwd" in kernels and "sos_warm_pair_bwd" in kernels, kernels
    seq = ops.sosfilt(ops.sosfilt(x, A, mode="warm"), B, mode="warm")
    scale = seq.abs().amax(dim=1)
    err = ((y - seq).abs().amax(dim=1) / scale).max().item()
    V = ops.pair_plan(C, T, True, A, B)[2]
    edge = ((y[:, :V] - seq[:, :V]).abs().amax(dim=1) / scale).max().item()
    print(f"pair vs sequential @ {fs} Hz: all {err:.2e}, left edge {edge:.2e} (V = {V})")
    assert err < 2e-6 and edge < 1e-6
    rows = [1, C - 2]
    xin = x[rows].cpu().numpy()
    n_ld = np.asarray(OS.filtfilt_pad(A.b, A.a, xin, dtype=np.longdouble, reference_edges=True), dtype=np.float64)
    truth = OS.filtfilt_pad(B.b, B.a, n_ld)
    ref = OS.filtfilt_pad(B.b, B.a, OS.filtfilt_pad(A.b, A.a, xin))
    err_gpu, err_ref = max_rel(y[rows].cpu().numpy(), truth), max_rel(ref, truth)
    print(f"pair vs long-double truth {err_gpu:.2e}; float64 reference vs truth {err_ref:.2e}")
    assert err_gpu <= max(TOL, err_ref) and err_gpu < TOL, (err_gpu, err_ref)     # exact numerator factors: 1e-5 outright


def test_fusion_groups_and_param_scope():
    from decode_tonal_langauge_b200 import preprocessor as P
    from decode_tonal_langauge_b200.chains import EX_STEPS, FULL6_STEPS
    kinds = [g[0] for g in P.fusion_groups(FULL6_STEPS)]
    assert kinds == ["iir_pair", "car_hilbert", "step", "step"]
    assert [g[0] for g in P.fusion_groups(EX_STEPS)] == ["step", "step", "step"]
    causal = [{"module": "preprocess.frequency_filter", "params": {"bands": [
        {"method": "butter", "params": {"freqs": [58, 62], "filter_type": "bandstop", "causal": True}}]}}] + FULL6_STEPS[1:]
    assert P.fusion_groups(causal)[0][0] == "step"          # a causal filter does not commute with a reversed sweep
    # the shared parameter Namespace: preserve_nans set on channel_zscore also governs a later
    # rolling_zscore (ref: preprocess/preprocessor.py:46-53), exclude_channels stays visible
    x = np.random.default_rng(0).standard_normal((4, 6000)).astype(np.float32)
    x[1] = 3.0                                                # constant row -> 0/0 in the z-score
    steps = [{"module": "preprocess.channel_zscore", "params": {"preserve_nans": False}},
             {"module": "preprocess.rolling_zscore", "params": {"window_length": 0.5}}]
    p = Namespace(signal_freq=1000)
    y, _ = P.preprocess_signal(x, steps, p)
    assert p.preserve_nans is False and np.all(y[:, 0] == 0) and np.isfinite(y).all()
    y2, _ = P.preprocess_signal(x, steps[1:], Namespace(signal_freq=1000))
    assert np.isnan(y2[:, 0]).all()


def test_fused_chain_small_matches_unfused_and_oracle():
    """CAR folded into the Hilbert load (and the pair's sequential fallback on a small recording),
    with excluded channels, against the step-by-step path and the oracle."""
    from oracle import chains as CH
    from decode_tonal_langauge_b200.chains import FULL6_STEPS
    from decode_tonal_langauge_b200.preprocessor import preprocess_signal
    from decode_tonal_langauge_b200 import synth
    fs, C, T = 2000, 12, 120_000
    x = synth.session_channels(2, range(C), T, fs, C)
    steps = [dict(s) for s in FULL6_STEPS]
    steps[1] = {"module": "preprocess.car_rereference", "params": {"exclude_channels": [2, 7]}}
    (y, f), kernels = _launched(lambda: preprocess_signal(x, steps, Namespace(signal_freq=fs)))
    assert "car_fused" not in kernels and "car_apply" not in kernels, kernels      # folded
    assert "car_colsum" in kernels and "hilbert_env8" in kernels
    y0, _ = preprocess_signal(x, steps, Namespace(signal_freq=fs), fuse=False)
    assert max_rel(y, y0) < 5e-6
    ref, fr = CH.run_chain(x, fs, steps)
    assert fr == f and max_rel(y, ref) < 5e-5               # the float64 notch of the reference: section 8c
    yf, _ = preprocess_signal(x, steps, Namespace(signal_freq=fs), output_dtype=np.float32)
    assert yf.dtype == np.float32 and y.dtype == np.float64 and max_rel(yf, y) < 1e-6


def test_strided_views_through_every_operator():
    """Row-padded views (big[:, :T], row stride > T) give what the dense copy gives, and nothing is
    written past the result (car / car_apply / zscore used x's stride for y)."""
    from decode_tonal_langauge_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(9)
    C, T, pad = 16, 24_000, 64
    big = torch.randn((C, T + pad), generator=g, device="cuda") * 10
    v, d = big[:, :T], big[:, :T].contiguous()
    assert v.stride(0) == T + pad
    guard = torch.full((C * T + 4096,), 7.0, device="cuda")            # results land in fresh allocations; cheap canary
    cs = ops.car_colsum(d)
    checks = {
        "car": lambda t: ops.car(t, [1, 5]),
        "car_apply": lambda t: ops.car_apply(t, cs, C),
        "car_colsum": lambda t: ops.car_colsum(t),
        "zscore": lambda t: ops.zscore(t, 100, 9000),
        "butter": lambda t: ops.butter(t, [70, 150], 2000.0),
        "butter_causal": lambda t: ops.butter(t, [70, 150], 2000.0, causal=True),
        "hilbert": lambda t: ops.hilbert(t, 2000.0, [70.0, 150.0]),
        "hilbert_car": lambda t: ops.hilbert(t, 2000.0, [70.0, 150.0], car=(cs, C)),
        "fir": lambda t: ops.fir_bank(t, 2000.0, 120, [80.0, 110.0]),
        "rolling": lambda t: ops.rolling_zscore(t, 500),
        "resample": lambda t: ops.fft_resample(t, T // 5),
        "resample_1stage": lambda t: ops.fft_resample(t, T // 5, two_stage=False),
    }
    for name, fn in checks.items():
        a, b = fn(v), fn(d)
        assert a.shape == b.shape, name
        assert torch.equal(torch.nan_to_num(a), torch.nan_to_num(b)), name
    assert torch.all(guard == 7.0)
    ref = d - d[[i for i in range(C) if i not in (1, 5)]].mean(dim=0, keepdim=True)
    assert ((ops.car(v, [1, 5]) - ref).abs().max() / ref.abs().max()).item() < 1e-6


def test_car_2048_rows_falls_back_to_two_phase():
    from decode_tonal_langauge_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn((2048, 8192), generator=g, device="cuda") * 5 + 1.0
    y, kernels = _launched(lambda: ops.car(x, [0, 2047]))
    assert "car_colsum" in kernels and "car_apply" in kernels, kernels
    ref = x - x[1:2047].mean(dim=0, keepdim=True, dtype=torch.float64).to(torch.float32)
    assert ((y - ref).abs().max() / ref.abs().max()).item() < 2e-6
    y2 = ops.car(x[:1500])                                    # still the fused strip
    ref2 = x[:1500] - x[:1500].mean(dim=0, keepdim=True, dtype=torch.float64).to(torch.float32)
    assert ((y2 - ref2).abs().max() / ref2.abs().max()).item() < 2e-6


def test_gather_small_dtypes_bit_exact():
    from decode_tonal_langauge_b200 import ops
    rng = np.random.default_rng(4)
    for dt in (np.int16, np.uint8, np.float16):
        src = rng.integers(0, 200, (5, 3000)).astype(dt)
        starts = np.array([0, 7, 1501, 2999 - 333], dtype=np.int64)
        out = ops.epoch_gather(torch.from_numpy(src).cuda(), starts, 333).cpu().numpy()
        assert out.dtype == dt and out.shape == (4, 5, 333)
        for n, s in enumerate(starts):
            assert np.array_equal(out[n], src[:, s:s + 333])
        ep = torch.from_numpy(out).cuda()
        assert np.array_equal(ops.channel_select(ep, [4, 0]).cpu().numpy(), out[:, [4, 0], :])
    with pytest.raises(ValueError, match="exceeds"):
        ops.epoch_gather(torch.zeros((2, 100), dtype=torch.int16, device="cuda"), np.array([90]), 20)


def test_anova_more_than_16_groups():
    """scipy's f_oneway has no group limit (discriminative.py:172-180 with a many-class target)."""
    from scipy import stats
    from decode_tonal_langauge_b200 import ops
    rng = np.random.default_rng(6)
    N, C, L, G = 900, 3, 40, 23
    labels = rng.integers(0, G, N)
    labels[:G] = np.arange(G)
    ep = (rng.standard_normal((N, C, L)) + 0.2 * labels[:, None, None] * (np.arange(L) > 20)).astype(np.float32)
    F, P = ops.anova_f(torch.from_numpy(ep).cuda(), labels)
    Fr, Pr = stats.f_oneway(*[ep[labels == k].astype(np.float64) for k in range(G)], axis=0)
    assert np.max(np.abs(F.cpu().numpy() - Fr) / np.abs(Fr)) < 1e-5
    ok = Pr > 1e-300
    assert np.max(np.abs(np.log10(P.cpu().numpy()[ok]) - np.log10(Pr[ok]))) < 1e-4


def test_table_cache_is_bounded(monkeypatch):
    from decode_tonal_langauge_b200 import ops
    monkeypatch.setattr(ops, "TABLE_CACHE_BYTES", 24 << 20)
    ops.release_workspaces()
    g = torch.Generator(device="cuda").manual_seed(1)
    for T in (240_000, 250_000, 270_000, 300_000, 320_000, 360_000):   # a session of blocks of different lengths
        x = torch.randn((2, T), generator=g, device="cuda")
        y = ops.fft_resample(x, T // 5, two_stage=False)
        assert y.shape == (2, T // 5)
        assert ops.table_cache_bytes() <= (24 << 20) + (8 << 20)
    ops.release_workspaces()
    assert ops.table_cache_bytes() == 0


def test_tma_sweeps_equal_cp_async_sweeps(monkeypatch):
    """csrc/sosfilt_tma.cu (cp.async.bulk.tensor boxes, SWIZZLE_128B, mbarriers) against the cp.async ring
    kernel: same algorithm, another chunk length and warm-up rounding -> equal to float32 rounding; single
    cascades in all three numerator forms and the fused pair, both row ends included."""
    from decode_tonal_langauge_b200 import design as D
    from decode_tonal_langauge_b200 import ops, synth
    fs, C, T = 2000, 128, 2_400_000
    x = synth.device_session(C, T, fs, seed=3)
    monkeypatch.setenv("ECOG_SOS_TMA", "0")

    def rel(a, b):
        return ((a - b).abs().amax(dim=1) / b.abs().amax(dim=1)).max().item()

    for freqs, ft in (([58, 62], "bandstop"), ([70, 150], "bandpass"), ([95, 105], "bandstop")):
        d = D.butter_design(freqs, fs, 4, False, ft)
        ref = ops.sosfilt(x, d, mode="warm")
        got, kernels = _launched(lambda: ops.sosfilt(x, d, mode="tma"))
        assert "sos_warm_tma_fwd" in kernels and "sos_warm_tma_bwd" in kernels, kernels
        e = rel(got, ref)
        print(f"TMA vs cp.async {ft} {freqs}: {e:.2e}")
        assert e < 1e-6
    A = D.butter_design([58, 62], fs, 4, False, "bandstop")
    B = D.butter_design([70, 150], fs, 4, False, "bandpass")
    # a channel chunk of the pipelined host path: 21 rows, the last CTA partly empty
    d = D.butter_design([58, 62], fs, 4, False, "bandstop")
    xs = x[:21]
    got, kernels = _launched(lambda: ops.sosfilt(xs, d, mode="tma"))
    assert "sos_warm_tma_fwd" in kernels and rel(got, ops.sosfilt(xs, d, mode="scan")) < 1e-6
    for C2, T2 in ((256, 4_800_000), (96, 9_600_000)):
        x2 = synth.device_session(C2, T2, fs, seed=4)
        monkeypatch.setenv("ECOG_SOS_TMA", "0")
        ref, k0 = _launched(lambda: ops.sosfilt_pair(x2, A, B))
        monkeypatch.setenv("ECOG_SOS_TMA", "1")
        monkeypatch.setenv("ECOG_PAIR_F32", "0")                    # all-float64 pair: the same arithmetic as the ring kernel
        got, kernels = _launched(lambda: ops.sosfilt_pair(x2, A, B))
        assert "sos_warm_pair_fwd" in k0 and "sos_warm_tma_pair_fwd" in kernels and "sos_warm_pair_fwd" not in kernels, (k0, kernels)
        e = rel(got, ref)
        print(f"TMA pair vs cp.async pair ({C2} x {T2}): {e:.2e}")
        assert e < 1e-6
        monkeypatch.delenv("ECOG_PAIR_F32")
        del x2, ref, got


@pytest.mark.parametrize("fs,C,T", [(2000, 256, 4_800_000), (3000, 128, 10_800_000)])
def test_pair_float32_bandpass_half(fs, C, T, monkeypatch):
    """ECOG_SOS_SPLIT_F32B (the default of the TMA pair when design.bandpass_f32_ok): band-pass sections in
    float32 delta form against the all-float64 pair on every row, and against the long-double truth on two rows.
    CPU emulation of the same recursion: 3.5e-7 (2 kHz) / 4.5e-7 (3 kHz)."""
    from oracle import steps as OS
    from decode_tonal_langauge_b200 import design as D
    from decode_tonal_langauge_b200 import ops, synth
    x = synth.device_session(C, T, fs, seed=9)
    A = D.butter_design([58, 62], fs, 4, False, "bandstop")
    B = D.butter_design([70, 150], fs, 4, False, "bandpass")
    assert D.bandpass_f32_ok(B) and not D.bandpass_f32_ok(A)
    assert not D.bandpass_f32_ok(D.butter_design([70, 150], 6000, 4, False, "bandpass"))      # poles too close to the circle
    monkeypatch.setenv("ECOG_PAIR_F32", "0")
    y64, k64 = _launched(lambda: ops.sosfilt_pair(x, A, B))
    monkeypatch.setenv("ECOG_PAIR_F32", "1")
    y32, k32 = _launched(lambda: ops.sosfilt_pair(x, A, B))
    assert "sos_warm_tma_pair_fwd" in k64 and "sos_pair_ws_fwd" in k32 and "sos_pair_ws_bwd" in k32, (k64, k32)
    e = ((y32 - y64).abs().amax(dim=1) / y64.abs().amax(dim=1)).max().item()
    rows = [0, C - 1]
    xin = x[rows].cpu().numpy()
    n_ld = np.asarray(OS.filtfilt_pad(A.b, A.a, xin, dtype=np.longdouble, reference_edges=True), dtype=np.float64)
    truth = OS.filtfilt_pad(B.b, B.a, n_ld)
    e32, e64 = max_rel(y32[rows].cpu().numpy(), truth), max_rel(y64[rows].cpu().numpy(), truth)
    print(f"pair @ {fs} Hz: float32 band-pass half vs float64 pair {e:.2e}; vs long-double truth {e32:.2e} (float64 pair {e64:.2e})")
    assert e < 1.5e-6 and e32 < 1.5e-6


def test_session_batch_runner_equals_one_by_one(tmp_path):
    """sessions.preprocess_sessions (file -> pinned -> device -> chain -> pinned -> sink, overlapped) gives what
    preprocess_signal gives block by block; the sink writes the reference's B<block>_ecog.npz files."""
    from decode_tonal_langauge_b200 import sessions as SS
    from decode_tonal_langauge_b200 import synth
    from decode_tonal_langauge_b200.chains import EX_STEPS, FULL6_STEPS
    from decode_tonal_langauge_b200.preprocessor import preprocess_signal
    fs = 2000
    paths, arrays = [], []
    for b, (C, T) in enumerate([(6, 60_000), (4, 90_000), (6, 60_000)]):
        x = synth.session_channels(b, range(C), T, fs, C)
        np.savez(tmp_path / f"B{b}_raw.npz", data=x, sf=np.float64(fs))
        paths.append(str(tmp_path / f"B{b}_raw.npz"))
        arrays.append(x)
    for steps in (FULL6_STEPS, EX_STEPS):
        sink = SS.save_block_sink(str(tmp_path / "out"), 1, [7, 8, 9])
        tm = {}
        written = SS.preprocess_sessions(paths, steps, sink, depth=2, timing=tm)
        assert tm["sessions"] == 3 and [p.endswith(f"B{k}_ecog.npz") for p, k in zip(written, (7, 8, 9))] == [True] * 3
        for p, x in zip(written, arrays):
            z = np.load(p)
            want, f = preprocess_signal(x, steps, Namespace(signal_freq=fs))
            assert z["sf"][()] == f and z["data"].dtype == want.dtype
            assert max_rel(z["data"], want) < 2e-6
    got = SS.preprocess_sessions([(arrays[1], fs)], FULL6_STEPS, lambda i, y, f: y.copy(), output_dtype=np.float32)
    assert got[0].dtype == np.float32
    with pytest.raises(ValueError):
        SS.preprocess_sessions([(arrays[0][0], fs)], FULL6_STEPS)
