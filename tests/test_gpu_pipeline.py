"""Plug-in level parity on the B200: step ``run(data, params)`` modules, the step runner
(chains EX and FULL6), epoch extraction and channel selection against the golden fixtures
written by the real reference."""
from argparse import Namespace

import numpy as np
import pytest

from conftest import max_rel

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module", autouse=True)
def _need_cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def test_step_modules_numpy_contract(golden):
    from decode_tonal_langauge_b200 import steps as S
    g = golden("steps")
    x, fs = g["x"], int(g["fs"])
    p = Namespace(signal_freq=fs, downsample_freq=400)
    y = S.downsample(x, p)
    assert isinstance(y, np.ndarray) and y.dtype == np.float32 and p.signal_freq == 400
    assert max_rel(y, g["downsample"]) < TOL
    p = Namespace(signal_freq=fs, bands=[
        {"method": "hilbert", "params": {"freq_ranges": [70.0, 150.0], "envelope": True}},
        {"method": "butter", "params": {"freqs": [0.3, 100], "filter_type": "bandpass"}}])
    y = S.frequency_filter(x, p)
    assert y.dtype == np.float64 and y.shape == g["two_bands"].shape
    C = x.shape[0]
    assert max_rel(y[:C], g["two_bands"][:C]) < TOL
    # 0.3-100 Hz at 2 kHz puts poles at radius 0.9996: the reference's float64 direct form is
    # 0.9 % away from the extended-precision evaluation of its own algorithm (section 8c rule)
    from oracle import steps as OS
    from decode_tonal_langauge_b200 import design as D
    d = D.butter_design([0.3, 100], fs, 4, False, "bandpass")
    truth = OS.filtfilt_pad(d.b, d.a, x, dtype=np.longdouble)
    err_gpu, err_ref = max_rel(y[C:], truth), max_rel(g["two_bands"][C:], truth)
    assert err_gpu <= max(TOL, err_ref) and err_gpu < TOL, (err_gpu, err_ref)
    p = Namespace(signal_freq=fs)
    y = S.car_rereference(x, p)
    assert p.exclude_channels == [] and y.dtype == np.float32 and max_rel(y, g["car"]) < 2e-6
    assert max_rel(S.channel_zscore(x, Namespace(signal_freq=fs)), g["channel_zscore"]) < 2e-6
    y = S.zscore_rereference(x, Namespace(signal_freq=fs, rereference_interval=[0.5, 3.0]))
    assert max_rel(y, g["zscore_rereference"]) < 2e-6
    # float64 input is accepted like the reference and keeps its dtype where the reference does
    assert S.car_rereference(x.astype(np.float64), Namespace(signal_freq=fs)).dtype == np.float64
    # the two secondary operators
    y = S.frequency_filter(x, Namespace(signal_freq=fs, bands=[
        {"method": "fir", "params": {"order": 390, "center_frequencies": [80.0, 100.0, 120.0]}}]))
    assert y.dtype == np.float32 and max_rel(y, g["fir"]) < TOL          # an all-fir list keeps the input dtype
    y = S.rolling_zscore(x, Namespace(signal_freq=fs, window_length=1.5))
    assert y.dtype == np.float64 and np.isnan(y[:, 0]).all()
    assert max_rel(y[:, 1:], g["rolling_zscore"][:, 1:]) < TOL
    y = S.rolling_zscore(x, Namespace(signal_freq=fs, window_length=1.5, preserve_nans=False))
    assert (y[:, 0] == 0).all()
    with pytest.raises(ValueError):
        S.rolling_zscore(x, Namespace(signal_freq=10, window_length=0.1))


def test_step_errors_match_reference_types(golden):
    from decode_tonal_langauge_b200 import steps as S
    x = golden("steps")["x"]
    with pytest.raises(ValueError):
        S.frequency_filter(x, Namespace(signal_freq=2000))
    with pytest.raises(ValueError):
        S.frequency_filter(x, Namespace(signal_freq=2000, bands=[{"method": "hilbert", "params": {}}]))
    with pytest.raises(ValueError):
        S.frequency_filter(x, Namespace(signal_freq=2000, bands=[{"method": "butter", "params": {}}]))
    with pytest.raises(ValueError):
        S.car_rereference(x, Namespace(exclude_channels=(1,)))
    with pytest.raises(ValueError):
        S.car_rereference(x, Namespace(exclude_channels=[99]))
    with pytest.raises(ValueError):
        S.zscore_rereference(x, Namespace(signal_freq=2000))
    with pytest.raises(ValueError):
        S.zscore_rereference(x, Namespace(signal_freq=2000, rereference_interval=[2.0, 1.0]))
    with pytest.raises(ValueError):
        S.zscore_rereference(x, Namespace(signal_freq=2000, rereference_interval=[0.0, 100.0]))


def test_chains_against_reference(golden):
    from decode_tonal_langauge_b200.chains import EX_STEPS, FULL6_STEPS
    from decode_tonal_langauge_b200.preprocessor import preprocess_signal
    c = golden("chains")
    x, fs = c["x"], int(c["fs"])
    p = Namespace(signal_freq=fs)
    y, f = preprocess_signal(x, FULL6_STEPS, p)
    assert f == int(c["full6_fs"]) and y.dtype == c["full6"].dtype and y.shape == c["full6"].shape
    err = max_rel(y, c["full6"])
    # the notch stage of this chain is the ill-conditioned design of section 8c: the float64
    # reference carries ~2e-5 of its own round-off into the chain output
    assert err < 5e-5, err
    p = Namespace(signal_freq=fs)
    y, f = preprocess_signal(x, EX_STEPS, p)
    assert f == int(c["ex_fs"]) and y.shape == c["ex"].shape and y.dtype == c["ex"].dtype
    assert max_rel(y, c["ex"]) < TOL
    # device-resident caller: tensors in, tensors out
    yd, _ = preprocess_signal(torch.from_numpy(x).cuda(), EX_STEPS, Namespace(signal_freq=fs))
    assert yd.is_cuda and max_rel(yd.cpu().numpy(), c["ex"]) < TOL


def test_chain_full6_long_double_rule(golden):
    """FULL6 against the extended-precision evaluation of the reference's own algorithm."""
    from oracle import steps as S
    from oracle import chains as CH
    from decode_tonal_langauge_b200 import design as D
    from decode_tonal_langauge_b200.chains import FULL6_STEPS
    from decode_tonal_langauge_b200.preprocessor import preprocess_signal
    c = golden("chains")
    x, fs = c["x"], int(c["fs"])
    d = D.butter_design([58, 62], fs, 4, False, "bandstop")
    notch_ld = np.asarray(S.filtfilt_pad(d.b, d.a, x, dtype=np.longdouble), dtype=np.float64)
    truth, _ = CH.run_chain(notch_ld, fs, FULL6_STEPS[1:])
    y, _ = preprocess_signal(x, FULL6_STEPS, Namespace(signal_freq=fs))
    err_gpu, err_ref = max_rel(y, truth), max_rel(c["full6"], truth)
    assert err_gpu <= max(TOL, err_ref), (err_gpu, err_ref)


@pytest.mark.parametrize("chain", ["FULL6", "nocar", "car_first"])
def test_pipelined_host_path_equals_monolithic(chain, monkeypatch):
    """Host arrays of >= 64 channels go through the channel-chunked, copy-overlapped path; it must
    give what the one-piece path gives (same kernels, different chunking) and match the oracle."""
    from decode_tonal_langauge_b200 import preprocessor as P
    from decode_tonal_langauge_b200.chains import FULL6_STEPS
    from oracle import chains as ochains
    steps = {"FULL6": FULL6_STEPS,
             "nocar": [FULL6_STEPS[4], FULL6_STEPS[2], FULL6_STEPS[5]],
             "car_first": [FULL6_STEPS[1], FULL6_STEPS[3], FULL6_STEPS[4]]}[chain]
    rng = np.random.default_rng(3)
    C, T, fs = 72, 40000, 2000
    common = np.cumsum(rng.standard_normal(T)) * 0.2
    x = (rng.standard_normal((C, T)) * 20 + common + 5 * np.sin(2 * np.pi * 60 * np.arange(T) / fs)).astype(np.float32)
    used = []
    real = P._pipelined_host_run
    monkeypatch.setattr(P, "_pipelined_host_run", lambda *a, **k: used.append(1) or real(*a, **k))
    y, f = P.preprocess_signal(x, steps, Namespace(signal_freq=fs))
    assert used, "the pipelined path was not taken"
    monkeypatch.setenv("ECOG_PIPELINE", "0")
    y0, f0 = P.preprocess_signal(x, steps, Namespace(signal_freq=fs))
    assert f == f0 and y.dtype == y0.dtype and y.shape == y0.shape
    assert max_rel(y, y0) < 2e-6
    ref, fr = ochains.run_chain(x, fs, steps)
    assert fr == f and max_rel(y, ref) < 5e-5


def test_strict_params_reproduces_reference_collision():
    from decode_tonal_langauge_b200.chains import FULL6_STEPS
    from decode_tonal_langauge_b200.preprocessor import preprocess_signal
    x = np.random.default_rng(0).standard_normal((2, 4000)).astype(np.float32)
    with pytest.raises(ValueError, match="already exists"):
        preprocess_signal(x, FULL6_STEPS, Namespace(signal_freq=2000), strict_params=True)


def test_epochs_against_reference(golden):
    from decode_tonal_langauge_b200 import epochs as E
    e = golden("epochs")
    intervals, rec = {}, {}
    for b in (1, 2):
        intervals[b] = {"start": e[f"b{b}_start"], "tone": e[f"b{b}_tone"],
                        "syllable": [str(s) for s in e[f"b{b}_syllable"]]}
        rec[b] = {"ecog": (e[f"b{b}_ecog"], e["ecog_sf"][()]), "audio": (e[f"b{b}_audio"], e["audio_sf"][()])}
    out = E.extract_epochs(intervals, rec, ["i", "a"], 1.0, (0.0, 5.0))
    order = [int(str(f)[1]) for f in e["listdir"] if "sound" in str(f)]
    n = len(e["b1_start"])
    perm = np.concatenate([np.arange(n) + n * (b - 1) for b in order])
    assert np.array_equal(out["ecog"][perm], e["ref_ecog"]) and out["ecog"].dtype == e["ref_ecog"].dtype
    assert np.array_equal(out["audio"][perm], e["ref_audio"])
    assert np.array_equal(out["syllable"][perm], e["ref_syllable"]) and out["syllable"].dtype == np.int8
    assert np.array_equal(out["tone"][perm], e["ref_tone"])
    assert out["ecog_rest"].shape == e["ref_ecog_rest"].shape
    assert np.array_equal(np.sort(out["ecog_rest"].ravel()), np.sort(e["ref_ecog_rest"].ravel()))
    with pytest.raises(ValueError, match="exceeds"):
        E.extract_epochs({1: {"start": np.array([39.5]), "tone": np.array([1]), "syllable": ["i"]}},
                         {1: rec[1]}, ["i", "a"], 1.0, None)
    with pytest.raises(ValueError, match="Mismatch"):
        E.extract_epochs(intervals, {1: rec[1], 2: {"ecog": rec[2]["ecog"]}}, ["i", "a"])


def test_extract_ecog_audio_files(golden, tmp_path):
    from decode_tonal_langauge_b200 import epochs as E
    e = golden("epochs")
    intervals = {}
    for b in (1, 2):
        np.savez(tmp_path / f"B{b}_ecog.npz", data=e[f"b{b}_ecog"], sf=e["ecog_sf"])
        np.savez(tmp_path / f"B{b}_sound.npz", data=e[f"b{b}_audio"], sf=e["audio_sf"])
        intervals[b] = {"start": e[f"b{b}_start"], "tone": e[f"b{b}_tone"],
                        "syllable": [str(s) for s in e[f"b{b}_syllable"]]}
    out_path = tmp_path / "subject_1.npz"
    out = E.extract_ecog_audio(intervals, str(tmp_path), ["i", "a"], 1.0, str(out_path), (0.0, 5.0))
    saved = np.load(out_path)
    assert set(saved.keys()) == {"ecog", "ecog_sf", "audio", "audio_sf", "syllable", "tone", "ecog_rest"}
    assert np.array_equal(saved["ecog"], out["ecog"]) and saved["ecog"].shape == e["ref_ecog"].shape
    assert np.array_equal(np.sort(saved["ecog"].ravel()), np.sort(e["ref_ecog"].ravel()))


def test_selection_against_reference(golden):
    from decode_tonal_langauge_b200 import selection as SEL
    s = golden("selection")
    data = {k: s[k] for k in ("ecog", "ecog_rest", "ecog_sf", "tone", "syllable")}
    for target in ("tone", "syllable"):
        for key in ("target", "label"):                          # Appendix B3: both spellings
            r = SEL.discriminative_run(data, {"p_threshold": 0.01, "active_time_threshold": 0.1, key: target})
            assert r["selected_channels"] == s[f"disc_{target}_selected"].tolist()
            assert all(type(c) is int for c in r["selected_channels"]) and r["max_lengths"] == []
        P, Pr = r["p_values"], s[f"disc_{target}_p"]
        assert np.array_equal(np.isnan(P), np.isnan(Pr))
        ok = np.isfinite(Pr) & (Pr > 1e-300)
        assert np.max(np.abs(np.log10(P[ok]) - np.log10(Pr[ok]))) < 1e-4
    r = SEL.active_run(data, {"p_threshold": 0.01, "active_time_threshold": 0.1})
    assert r["selected_channels"] == s["active_selected"].tolist()
    assert r["max_lengths"] == s["active_max_lengths"].tolist()
    ok = s["active_p_last"] > 1e-300
    assert np.max(np.abs(np.log10(r["p_values"][ok]) - np.log10(s["active_p_last"][ok]))) < 1e-4
    with pytest.raises(KeyError):
        SEL.discriminative_run(data, {"active_time_threshold": 0.1, "target": "nope"})
    with pytest.raises(ValueError):
        SEL.discriminative_run({**data, "tone": data["tone"].astype(np.float64)},
                               {"active_time_threshold": 0.1, "target": "tone"})


def test_full_size_properties():
    """Size-independent properties at a BASELINE-scale row count (one 60-min row would not
    finish on the CPU oracle in test time): linearity of the linear steps, CAR idempotence,
    z-score moments, resample of a band-limited tone."""
    from decode_tonal_langauge_b200 import ops
    C, T, fs = 8, 7_200_000, 2000.0
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.randn((C, T), generator=g, device="cuda") * 30
    b = torch.randn((C, T), generator=g, device="cuda") * 30
    fa, fb = ops.butter(a, [70, 150], fs), ops.butter(b, [70, 150], fs)
    fab = ops.butter(a + 2 * b, [70, 150], fs)
    scale = fab.abs().max().item()
    assert (fab - (fa + 2 * fb)).abs().max().item() / scale < 2e-6
    ca = ops.car(a)
    assert ca.mean(dim=0).abs().max().item() < 1e-4
    assert (ops.car(ca) - ca).abs().max().item() < 1e-4
    z = ops.zscore(fa)
    assert z.mean(dim=1, dtype=torch.float64).abs().max().item() < 1e-6
    assert (z.to(torch.float64).std(dim=1, unbiased=False) - 1).abs().max().item() < 1e-6
    t = torch.arange(T, device="cuda", dtype=torch.float64) / fs
    tone = torch.sin(2 * np.pi * 37.0 * t).to(torch.float32)[None].repeat(2, 1)     # periodic over the row
    y = ops.fft_resample(tone, T // 5)
    t2 = torch.arange(T // 5, device="cuda", dtype=torch.float64) / 400.0
    assert (y[0].to(torch.float64) - torch.sin(2 * np.pi * 37.0 * t2)).abs().max().item() < 1e-5
    env = ops.hilbert(tone, fs, [30.0, 45.0])
    assert env.shape == tone.shape and torch.isfinite(env).all()
    # envelope of a pure in-band tone (periodic over the row) is constant: mean_b G_b(f) (A1)
    from decode_tonal_langauge_b200 import design as D
    f_tone = 100.0
    tone = torch.sin(2 * np.pi * f_tone * t).to(torch.float32)[None].repeat(2, 1)
    env = ops.hilbert(tone, fs, [70.0, 150.0])
    cfs, sds = D.gaussian_bank([70.0, 150.0])
    want = float(np.mean(np.exp(-0.5 * ((f_tone - cfs) / sds) ** 2)))
    assert (env.to(torch.float64) - want).abs().max().item() / want < 1e-5
    # two-stage brick wall == whole-row FFT on white noise at full length
    one = ops.fft_resample(a[:4], T // 5, two_stage=False)
    two = ops.fft_resample(a[:4], T // 5, two_stage=True)
    assert ((one - two).abs().amax(dim=1) / one.abs().amax(dim=1)).max().item() < 3e-6
    # notch at full length: warm-up path against the exact carry scan
    n1 = ops.butter(a[:4], [58, 62], fs, filter_type="bandstop", mode="warm")
    n2 = ops.butter(a[:4], [58, 62], fs, filter_type="bandstop", mode="scan")
    assert ((n1 - n2).abs().amax(dim=1) / n2.abs().amax(dim=1)).max().item() < 1e-6


def test_full_size_gather_and_selection_properties():
    """BASELINE configs[2] scale: 20k events x 256 ch x 400 samples.  Gather is a bit copy (random
    rows checked against slicing, plus a whole-tensor checksum against the analytic multiplicity);
    F is invariant to shifting / scaling the data and to relabelling the groups."""
    from decode_tonal_langauge_b200 import ops
    C, T, N, L = 256, 1_440_000, 20_000, 400
    g = torch.Generator(device="cuda").manual_seed(3)
    src = torch.randn((C, T), generator=g, device="cuda")
    rng = np.random.default_rng(3)
    starts = np.sort(rng.integers(0, T - L, N)).astype(np.int64)
    ep = ops.epoch_gather(src, starts, L)
    assert ep.shape == (N, C, L)
    for n in rng.integers(0, N, 50):
        c = int(rng.integers(0, C))
        assert torch.equal(ep[n, c], src[c, starts[n]:starts[n] + L])
    cover = np.zeros(T + 1)
    np.add.at(cover, starts, 1.0)
    np.add.at(cover, starts + L, -1.0)
    mult = torch.from_numpy(np.cumsum(cover)[:T]).cuda()
    total = (src.to(torch.float64) * mult[None]).sum().item()
    assert abs(ep.sum(dtype=torch.float64).item() - total) < 1e-6 * max(1.0, abs(total)) + 1e-3
    labels = rng.integers(0, 4, N)
    F, P = ops.anova_f(ep, labels)
    F2, _ = ops.anova_f(ep * 3.0 + 7.0, (labels + 1) % 4)
    ok = torch.isfinite(F)
    assert ((F - F2).abs()[ok] / F.abs().clamp_min(1e-3)[ok]).max().item() < 1e-4
    assert ((P >= 0) & (P <= 1)).all()
    runs = ops.sig_runlength(P, 0.01 / L)
    assert runs.shape == (C,) and int(runs.max()) <= L


def test_sample_handler_matches_reference_semantics(tmp_path):
    """ref: data_loading/sample_loading.py:34-123: channel union from the selection JSON, joint label
    code, features[:, channels, :] -- here gathered on the device, bit exact."""
    import json
    from decode_tonal_langauge_b200.samples import ClassificationSampleHandler
    from decode_tonal_langauge_b200 import ops
    rng = np.random.default_rng(0)
    N, C, L = 37, 19, 50
    ecog = rng.standard_normal((N, C, L))
    tone = rng.integers(0, 4, N)
    syl = rng.integers(0, 2, N).astype(np.int8)
    np.savez(tmp_path / "subject_1.npz", ecog=ecog, tone=tone, syllable=syl, ecog_sf=np.int64(400))
    sel = {"tone_discriminative": [3, 7, 11], "syllable_discriminative": [7, 2], "active": [1]}
    with open(tmp_path / "subject_1.json", "w") as f:
        json.dump(sel, f)
    p = Namespace(sample_path=str(tmp_path / "subject_1.npz"), channel_file=str(tmp_path / "subject_1.json"),
                  targets=["tone", "syllable"], features="ecog")
    d = ClassificationSampleHandler(p).load_data()
    assert d["selected_channels"].tolist() == [2, 3, 7, 11]
    assert d["features"].is_cuda and d["features"].dtype == torch.float64
    assert np.array_equal(d["features"].cpu().numpy(), ecog[:, [2, 3, 7, 11], :])
    assert np.array_equal(d["labels"], tone + 4 * syl.astype(int)) and d["n_classes_dict"] == {"tone": 4, "syllable": 2}
    h = ClassificationSampleHandler(Namespace(sample_path=p.sample_path, targets="tone", features="ecog"))
    d = h.load_data(as_numpy=True)
    assert np.array_equal(d["features"], ecog) and d["selected_channels"].tolist() == list(range(C))
    ds = h.prepare_torch_dataset(d["features"], d["labels"], "cuda")
    assert ds.tensors[0].dtype == torch.float32 and ds.tensors[0].shape == (N, C, L)
    with pytest.raises(KeyError):
        ClassificationSampleHandler(Namespace(sample_path=p.sample_path, channel_file=p.channel_file,
                                              targets=["nope"], features="ecog")).load_data()
    with pytest.raises(IndexError):
        ops.channel_select(torch.zeros((2, 3, 4), device="cuda"), [5])
    x32 = torch.randn((5, 6, 7), device="cuda")
    assert torch.equal(ops.channel_select(x32, [5, 0, -1]), x32[:, [5, 0, 5], :])
