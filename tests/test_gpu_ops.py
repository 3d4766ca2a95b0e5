"""Kernel-level parity on the B200: every C-ABI operator against the oracle / the golden
fixtures written by the real reference.  Tolerances follow SURVEY.md section 8c:
per-channel max-norm relative error <= 1e-5 for float outputs, bit-exact for copies,
indices and selected sets."""
import numpy as np
import pytest

from conftest import max_rel

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

TOL = 1e-5


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from decode_tonal_langauge_b200 import ops as O
    return O


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.cpu().numpy()


# ------------------------------------------------------------------ K1 / K2
@pytest.mark.parametrize("C,T", [(3, 12000), (5, 1001), (256, 4096), (130, 777), (700, 256)])
def test_car_matches_oracle(ops, C, T):
    from oracle import steps as S
    rng = np.random.default_rng(C * T)
    x = (rng.standard_normal((C, T)) * 30 + rng.standard_normal((1, T)) * 10).astype(np.float32)
    y = host(ops.car(dev(x)))
    assert max_rel(y, S.car_rereference(x)) < 2e-6
    excl = [0, C - 1] if C > 2 else [0]
    y = host(ops.car(dev(x), excl))
    assert max_rel(y, S.car_rereference(x, excl)) < 2e-6
    # two-phase (channel-sharded) form must agree with the fused pass
    xs = dev(x)
    s = ops.car_colsum(xs[: C // 2 + 1]) + ops.car_colsum(xs[C // 2 + 1:]) if C > 2 else ops.car_colsum(xs)
    y2 = host(ops.car_apply(xs, s, C))
    assert max_rel(y2, S.car_rereference(x)) < 2e-6


def test_car_golden_and_errors(ops, golden):
    g = golden("steps")
    assert max_rel(host(ops.car(dev(g["x"]))), g["car"]) < 2e-6
    assert max_rel(host(ops.car(dev(g["x"]), [1])), g["car_excl"]) < 2e-6
    with pytest.raises(ValueError):
        ops.car(dev(g["x"]), [7])
    with pytest.raises(ValueError):
        ops.car(dev(g["x"]), (1,) if False else "1")


def test_zscores_golden(ops, golden):
    g = golden("steps")
    x, fs = g["x"], int(g["fs"])
    assert max_rel(host(ops.zscore(dev(x))), g["channel_zscore"]) < 2e-6
    y = host(ops.zscore(dev(x), int(0.5 * fs), int(3.0 * fs)))
    assert max_rel(y, g["zscore_rereference"]) < 2e-6
    with pytest.raises(ValueError):
        ops.zscore(dev(x), 10, 10)
    with pytest.raises(ValueError):
        ops.zscore(dev(x), 0, x.shape[1] + 1)


@pytest.mark.parametrize("T", [999, 70000, 200001])
def test_row_stats_large_offset(ops, T):
    rng = np.random.default_rng(T)
    x = (rng.standard_normal((4, T)) * 0.01 + 5000.0).astype(np.float32)     # mean >> std
    mean, std = ops.row_stats(dev(x))
    x64 = x.astype(np.float64)
    assert np.allclose(host(mean), x64.mean(axis=1), rtol=1e-12)
    assert np.allclose(host(std), x64.std(axis=1), rtol=1e-9)
    const = np.full((2, T), 3.0, dtype=np.float32)
    z = host(ops.zscore(dev(const)))
    assert np.isnan(z).all()                      # 0/0 like the reference
    assert (host(ops.zscore(dev(const), nan_to_zero=True)) == 0).all()


# ------------------------------------------------------------------------ K3
@pytest.mark.parametrize("key,freqs,btype", [
    ("bandpass", [70, 150], "bandpass"), ("lowpass", 200.0, "lowpass"), ("highpass", 1.0, "highpass")])
@pytest.mark.parametrize("chunk", [None, 1024, 4096])
@pytest.mark.parametrize("mode", ["scan", "warm"])
def test_butter_filtfilt_golden(ops, golden, key, freqs, btype, chunk, mode):
    g = golden("steps")
    if mode == "warm" and key == "highpass":
        # a 1 Hz high-pass remembers its state for longer than this row: the warm-up path declines
        with pytest.raises(ValueError):
            ops.butter(dev(g["x"]), freqs, float(g["fs"]), 4, False, btype, chunk=chunk, mode=mode)
        return
    y = host(ops.butter(dev(g["x"]), freqs, float(g["fs"]), 4, False, btype, chunk=chunk, mode=mode))
    assert max_rel(y, g[key]) < TOL


@pytest.mark.parametrize("chunk", [None, 1024, 2000 * 16 // 16 * 16])
@pytest.mark.parametrize("mode", ["scan", "warm"])
def test_notch_long_double_rule(ops, golden, chunk, mode):
    """4 Hz band-stop: the float64 reference itself is ~2e-5 from the extended-precision
    evaluation of its own algorithm, so parity is stated against that (section 8c)."""
    from oracle import steps as S
    from decode_tonal_langauge_b200 import design as D
    g = golden("steps")
    x, fs = g["x"], float(g["fs"])
    d = D.butter_design([58, 62], fs, 4, False, "bandstop")
    y = host(ops.sosfilt(dev(x), d, chunk, mode=mode))
    truth = S.filtfilt_pad(d.b, d.a, x, dtype=np.longdouble)
    err_gpu, err_ref = max_rel(y, truth), max_rel(g["notch"], truth)
    assert err_gpu <= max(TOL, err_ref), (err_gpu, err_ref)
    assert max_rel(y, g["notch"]) < 5e-5


@pytest.mark.parametrize("mode", ["scan", "warm"])
def test_causal_sosfilt_golden(ops, golden, mode):
    g = golden("steps")
    y = host(ops.butter(dev(g["x"]), [70, 150], float(g["fs"]), 4, True, "bandpass", chunk=1024, mode=mode))
    assert max_rel(y, g["causal"]) < TOL


@pytest.mark.parametrize("C,T,chunk", [(2, 30, 1024), (3, 4097, 1024), (7, 33333, 2048), (300, 5000, 1024),
                                       (4, 40000, 2048), (600, 20000, 2048)])
@pytest.mark.parametrize("mode", ["scan", "warm"])
def test_filtfilt_ragged_shapes(ops, C, T, chunk, mode):
    from oracle import steps as S
    rng = np.random.default_rng(T)
    x = np.cumsum(rng.standard_normal((C, T)), axis=1).astype(np.float32)
    y = host(ops.butter(dev(x), [70, 150], 2000.0, 4, False, "bandpass", chunk=chunk, mode=mode))
    assert max_rel(y, S.butter_filter(x, [70, 150], 2000.0)) < TOL


def test_filtfilt_notch_long_row_warm_vs_scan(ops):
    """Long rows, narrow notch: chunks far from the row edges start from a zero-state warm-up;
    the result must agree with the exact carry scan far below the parity tolerance."""
    from decode_tonal_langauge_b200 import design as D
    rng = np.random.default_rng(11)
    x = dev((rng.standard_normal((6, 400_000)) * 30).astype(np.float32))
    d = D.butter_design([58, 62], 2000.0, 4, False, "bandstop")
    a = host(ops.sosfilt(x, d, 32768, mode="warm"))
    b = host(ops.sosfilt(x, d, 4096, mode="scan"))
    assert max_rel(a, b) < 5e-7


def test_filtfilt_too_short_raises(ops):
    with pytest.raises(ValueError):
        ops.butter(dev(np.zeros((2, 27), np.float32)), [70, 150], 2000.0)


# ------------------------------------------------------------------------ K4
def test_hilbert_golden(ops, golden):
    g = golden("steps")
    x, fs = dev(g["x"]), float(g["fs"])
    assert max_rel(host(ops.hilbert(x, fs, [70.0, 150.0])), g["hilbert_env"]) < TOL
    assert max_rel(host(ops.hilbert(x, fs, [70, 150])), g["hilbert_env"]) < TOL           # int spelling (B2)
    assert max_rel(host(ops.hilbert(x, fs, [70.0, 150.0], envelope=False)), g["hilbert_real"]) < TOL
    y = host(ops.hilbert(x, fs, [[30.0, 55.0], [70.0, 150.0]]))
    assert max_rel(y, g["hilbert_two_ranges"]) < TOL


@pytest.mark.parametrize("C,T,fs", [(1, 3000, 400.0), (5, 50001, 2000.0), (2, 4096, 1000.0)])
def test_hilbert_shapes(ops, C, T, fs):
    from oracle import steps as S
    rng = np.random.default_rng(T)
    x = (rng.standard_normal((C, T)) * 20).astype(np.float32)
    y = host(ops.hilbert(dev(x), fs, [70.0, 150.0]))
    assert max_rel(y, S.hilbert_filter(x, fs, [70.0, 150.0])) < TOL


@pytest.mark.parametrize("fs,rng_hz", [(3000.0, [70.0, 150.0]), (2000.0, [80.0, 120.0]), (3000.0, [100.0, 140.0])])
def test_hilbert_pruned_windows(ops, fs, rng_hz):
    """Banks whose windows use fewer than 16 bin groups (3 kHz: clamped to 8) and odd band counts."""
    from oracle import steps as S
    rng = np.random.default_rng(17)
    x = (rng.standard_normal((3, 30011)) * 20).astype(np.float32)
    x += (5 * np.sin(2 * np.pi * 120.0 * np.arange(30011) / fs)).astype(np.float32)      # a line inside the bank
    y = host(ops.hilbert(dev(x), fs, rng_hz))
    assert max_rel(y, S.hilbert_filter(x, fs, rng_hz)) < TOL


@pytest.mark.parametrize("envelope", [True, False])
def test_hilbert_low_frequency_bands_whole_record(ops, golden, envelope):
    """Theta / alpha bands at 2 kHz: time kernels longer than the block -> whole-record FFT path."""
    from oracle import steps as S
    g = golden("steps")
    x, fs = g["x"], float(g["fs"])
    ref = S.hilbert_filter(x, fs, [[4.0, 8.0], [8.0, 13.0]], envelope=envelope)
    y = host(ops.hilbert(dev(x), fs, [[4.0, 8.0], [8.0, 13.0]], envelope=envelope))
    assert max_rel(y, ref) < TOL


@pytest.mark.parametrize("envelope", [True, False])
def test_hilbert_delta_band_any_row_length(ops, envelope):
    """Low-frequency bands need the whole-record transform (frequency_filter.py:154-184 accepts any T):
    smooth lengths through the four-step FFT, prime lengths (real TDT rates) through the chirp-z form."""
    from oracle import steps as S
    rng = np.random.default_rng(2)
    for T in (8000, 8009, 100003):
        x = (np.cumsum(rng.standard_normal((2, T)), axis=1)).astype(np.float32)
        y = host(ops.hilbert(dev(x), 400.0, [0.5, 4.0], envelope=envelope))
        assert max_rel(y, S.hilbert_filter(x, 400.0, [0.5, 4.0], envelope=envelope)) < TOL, T


# ------------------------------------------------------------------ K6 / K7
def test_fir_bank_golden(ops, golden):
    g = golden("steps")
    y = host(ops.fir_bank(dev(g["x"]), float(g["fs"]), 390, [80.0, 100.0, 120.0]))
    assert max_rel(y, g["fir"]) < TOL


@pytest.mark.parametrize("C,T,ntaps", [(3, 5000, 391), (2, 777, 5), (4, 20001, 1025), (1, 100, 300)])
def test_fir_causal_vs_lfilter(ops, C, T, ntaps):
    from scipy import signal as sp_signal
    rng = np.random.default_rng(ntaps)
    x = (rng.standard_normal((C, T)) * 10).astype(np.float32)
    h = rng.standard_normal(ntaps) / ntaps
    y = host(ops.fir_causal(dev(x), h))
    assert max_rel(y, sp_signal.lfilter(h, 1.0, x.astype(np.float64), axis=-1)) < 2e-6


def test_rolling_zscore_golden(ops, golden):
    g = golden("steps")
    fs = float(g["fs"])
    y = host(ops.rolling_zscore(dev(g["x"]), int(1.5 * fs)))
    ref = g["rolling_zscore"]
    assert np.isnan(y[:, 0]).all() and np.isnan(ref[:, 0]).all()
    assert max_rel(y[:, 1:], ref[:, 1:]) < TOL
    z = host(ops.rolling_zscore(dev(g["x"]), int(1.5 * fs), nan_to_zero=True))
    assert (z[:, 0] == 0).all() and np.array_equal(z[:, 1:], y[:, 1:])


@pytest.mark.parametrize("C,T,W", [(3, 4097, 2), (2, 70000, 4000), (5, 9001, 20000), (2, 33, 16), (1, 50000, 4099)])
def test_rolling_zscore_vs_oracle(ops, C, T, W):
    from oracle import steps as S
    rng = np.random.default_rng(W)
    x = (np.cumsum(rng.standard_normal((C, T)), axis=1) + 30 * rng.standard_normal((C, T)) + 100).astype(np.float32)
    y = host(ops.rolling_zscore(dev(x), W))
    ref = S.rolling_zscore(x, 1.0, W)              # fs = 1: window_length is in samples
    assert max_rel(y[:, 1:], ref[:, 1:]) < TOL


# ------------------------------------------------------------------------ K5
def test_resample_golden(ops, golden):
    g = golden("steps")
    assert max_rel(host(ops.fft_resample(dev(g["x"]), 2400)), g["downsample"]) < TOL
    assert max_rel(host(ops.fft_resample(dev(g["x"][:2]), 3600)), g["downsample_600"]) < TOL


@pytest.mark.parametrize("T,num", [(12000, 2400), (20000, 4000), (57600, 11520), (240000, 48000), (6000, 9000),
                                   (8000, 8000), (1_200_000, 240_000)])
def test_resample_vs_oracle(ops, T, num):
    from oracle import steps as S
    rng = np.random.default_rng(T)
    C = 3 if T < 1_000_000 else 2
    x = (np.cumsum(rng.standard_normal((C, T)), axis=1) * 0.3 + rng.standard_normal((C, T)) * 5).astype(np.float32)
    y = host(ops.fft_resample(dev(x), num))
    assert max_rel(y, S.fft_resample(x.astype(np.float64), num)) < TOL


@pytest.mark.parametrize("T,num", [(12000, 2400), (57600, 11520), (240000, 32000), (1_200_000, 240_000)])
def test_resample_two_stage_equals_single_fft(ops, T, num):
    """FIR pre-decimation + compensated brick wall against the whole-row FFT and the oracle,
    on white noise (worst case for aliasing into the kept band)."""
    from oracle import steps as S
    from decode_tonal_langauge_b200 import fftplan as FP
    assert FP.predecimation(T, num) is not None
    rng = np.random.default_rng(T + 1)
    x = (rng.standard_normal((2, T)) * 30).astype(np.float32)
    two = host(ops.fft_resample(dev(x), num, two_stage=True))
    one = host(ops.fft_resample(dev(x), num, two_stage=False))
    ref = S.fft_resample(x.astype(np.float64), num)
    assert max_rel(two, ref) < 3e-6 and max_rel(one, ref) < 3e-6


@pytest.mark.parametrize("T,D", [(12000, 4), (12000, 2), (4100, 2), (65536 + 8, 4)])
def test_fir_decimate_matches_model(ops, T, D):
    from helpers import emulate as EM
    from decode_tonal_langauge_b200 import fftplan as FP
    rng = np.random.default_rng(T)
    x = (rng.standard_normal((3, T)) * 10).astype(np.float32)
    for ntaps, off in ((160, 80), (28, 12), (7, 3), (256, 128)):
        taps = (rng.standard_normal(ntaps) / ntaps).astype(np.float32)
        pre = FP.PreDecimation(D, taps, off, np.ones(1, np.float32))
        y = host(ops.fir_decimate(dev(x), taps, off, D))
        ref = np.stack([EM.fir_decimate_model(r.astype(np.float64).astype(np.float32), pre) for r in x])
        ref64 = np.stack([np.array([np.dot(taps.astype(np.float64), r[(m * D + np.arange(ntaps) - off) % T])
                                    for m in range(T // D)]) for r in x.astype(np.float64)])
        assert y.shape == ref.shape and max_rel(y, ref64) < 2e-6


@pytest.mark.parametrize("T", [12000, 8192 * 4, 8192 * 4 + 12, 200_004, 1_200_000])
def test_halfband2_decimate_matches_model(ops, T):
    """ecog_halfband2_decimate (two circular half-band stages, intermediate row in shared memory) against the float64
    evaluation of the same two sums: arbitrary taps of every supported length, rows that end inside a CTA's tile,
    a strided view."""
    from helpers import emulate as EM
    rng = np.random.default_rng(T)
    big = (rng.standard_normal((3, T + 8)) * 10).astype(np.float32)
    x = big[:, :T]
    for k1, k2 in ((8, 21), (7, 20), (1, 1), (3, 12)):
        s1 = (rng.standard_normal(1 + k1) / (1 + 2 * k1)).astype(np.float32)
        s2 = (rng.standard_normal(1 + k2) / (1 + 2 * k2)).astype(np.float32)
        y = host(ops.halfband2_decimate(dev(big)[:, :T], s1, s2))
        ref = np.stack([EM.halfband_stage_model(EM.halfband_stage_model(r.astype(np.float64), s1.astype(np.float64)).astype(np.float64),
                                                s2.astype(np.float64)) for r in x])
        assert y.shape == (3, T // 4) and max_rel(y, ref) < 2e-6, (k1, k2)


def test_resample_odd_golden(ops, golden):
    """Odd row length: the reference's own output for a 9001-sample row (chirp-z path)."""
    g = golden("steps")
    assert max_rel(host(ops.fft_resample(dev(g["x_odd"]), 1800)), g["downsample_odd"]) < TOL


@pytest.mark.parametrize("T,num", [(9001, 1800), (18310, 2399), (7919, 1000), (1000, 1501), (4099, 4099),
                                   (12002, 2400), (183_106, 23_999)])
def test_resample_any_length_vs_oracle(ops, T, num):
    """Non-smooth / odd / prime lengths (real TDT rates) through Bluestein, with and without the
    FIR pre-decimation stage."""
    from oracle import steps as S
    rng = np.random.default_rng(T + num)
    x = (np.cumsum(rng.standard_normal((3, T)), axis=1) * 0.3 + rng.standard_normal((3, T)) * 5).astype(np.float32)
    ref = S.fft_resample(x.astype(np.float64), num)
    assert max_rel(host(ops.fft_resample(dev(x), num)), ref) < TOL
    assert max_rel(host(ops.fft_resample(dev(x), num, two_stage=False)), ref) < TOL


def test_fft_c2c_matches_numpy(ops):
    from decode_tonal_langauge_b200 import fftplan as FP
    rng = np.random.default_rng(1)
    for N in (360, 3600, 36000, 196608):
        z = (rng.standard_normal((2, N)) + 1j * rng.standard_normal((2, N))).astype(np.complex64)
        zd = torch.view_as_real(torch.from_numpy(z)).contiguous().cuda()
        bp = FP.big_plan(N, FP.MAX_AXIS_NARROW)
        Z = torch.view_as_complex(ops.fft_c2c(zd.clone(), bp)).cpu().numpy()
        ref = np.fft.fft(z.astype(np.complex128), axis=1)
        assert np.max(np.abs(Z - ref)) / np.max(np.abs(ref)) < 2e-6
        back = torch.view_as_complex(ops.fft_c2c(zd.clone(), bp, inverse=True, scale=1.0 / N)).cpu().numpy()
        ref = np.fft.ifft(z.astype(np.complex128), axis=1)
        assert np.max(np.abs(back - ref)) / np.max(np.abs(ref)) < 2e-6


# ------------------------------------------------------------------------ K8
def test_gather_bit_exact(ops, golden):
    from oracle import epochs as E
    e = golden("epochs")
    for b in (1, 2):
        src = e[f"b{b}_ecog"]
        first, n = E.onset_indices(e[f"b{b}_start"], e["ecog_sf"][()], 1.0)
        out = host(ops.epoch_gather(dev(src), first, n))
        assert np.array_equal(out, E.gather(src, first, n))
        aud = e[f"b{b}_audio"]
        first, n = E.onset_indices(e[f"b{b}_start"], e["audio_sf"][()], 1.0)
        out = host(ops.epoch_gather(dev(aud), first, n))
        assert np.array_equal(out, E.gather(aud, first, n))
    # float64 and int64 sources, odd lengths, unaligned starts
    rng = np.random.default_rng(1)
    for dtype in (np.float64, np.int64, np.int32, np.float32):
        src = (rng.standard_normal((5, 3001)) * 1000).astype(dtype)
        first = np.array([0, 1, 2, 3, 1234, 3001 - 77], dtype=np.int64)
        assert np.array_equal(host(ops.epoch_gather(dev(src), first, 77)), E.gather(src, first, 77))
    assert ops.epoch_gather(dev(src), np.zeros(0, np.int64), 10).shape == (0, 5, 10)


def test_gather_rejects_overrun(ops):
    src = dev(np.zeros((2, 100), np.float32))
    with pytest.raises(ValueError):
        ops.epoch_gather(src, np.array([50]), 60)
    with pytest.raises(ValueError):
        ops.epoch_gather(src, np.array([-1]), 10)


# -------------------------------------------------------------------- K9 / K10
def test_anova_golden(ops, golden):
    s = golden("selection")
    ecog = dev(s["ecog"].astype(np.float32))
    for target in ("tone", "syllable"):
        labels = s[target]
        uniq, inv = np.unique(labels, return_inverse=True)
        F, P = ops.anova_f(ecog, inv)
        F, P = host(F), host(P)
        Fr, Pr = s[f"disc_{target}_f"], s[f"disc_{target}_p"]
        assert np.array_equal(np.isnan(F), np.isnan(Fr))
        ok = np.isfinite(Fr)
        assert np.max(np.abs(F[ok] - Fr[ok]) / np.abs(Fr[ok])) < TOL
        okp = ok & (Pr > 1e-300)
        assert np.max(np.abs(np.log10(P[okp]) - np.log10(Pr[okp]))) < 1e-4
        assert np.array_equal(np.isnan(P), np.isnan(Pr))
        runs = host(ops.sig_runlength(dev(P), 0.01 / P.shape[1]))
        sel = [int(c) for c in np.nonzero(runs > int(0.1 * 100))[0]]
        assert sel == s[f"disc_{target}_selected"].tolist()


def test_anova_two_tensors_and_edge_cases(ops):
    from oracle import selection as SEL
    rng = np.random.default_rng(3)
    erp = rng.standard_normal((37, 6, 50)).astype(np.float32) + 2
    rest = rng.standard_normal((9, 6, 50)).astype(np.float32) + 2
    erp[:, 1, 10:30] += 3
    erp[:, 2, :] = 1.5
    rest[:, 2, :] = 1.5                 # everything identical -> NaN
    erp[:, 3, :] = 2.0
    rest[:, 3, :] = 1.0                 # constant within groups, different between -> inf
    groups = np.r_[np.ones(37, np.int32), np.zeros(9, np.int32)]
    F, P = ops.anova_f(dev(erp), groups, dev(rest))
    F, P = host(F), host(P)
    for ch in range(6):
        Fr, Pr = SEL.anova_f([rest[:, ch, :].astype(np.float64), erp[:, ch, :].astype(np.float64)])
        assert np.array_equal(np.isnan(F[ch]), np.isnan(Fr))
        assert np.array_equal(np.isinf(F[ch]), np.isinf(Fr))
        ok = np.isfinite(Fr)
        if ok.any():
            assert np.max(np.abs(F[ch][ok] - Fr[ok]) / np.abs(Fr[ok])) < TOL
            assert np.allclose(P[ch][ok], Pr[ok], rtol=1e-6, atol=1e-300)
    assert (P[3] == 0).all() and np.isnan(P[2]).all()


def test_runlength_bit_exact(ops):
    from oracle import selection as SEL
    rng = np.random.default_rng(5)
    for L in (1, 31, 32, 33, 100, 400, 1000):
        p = rng.random((9, L))
        p[0] = 0.0
        p[1] = 1.0
        p[2, ::2] = np.nan
        thr = 0.4
        got = host(ops.sig_runlength(dev(p), thr))
        for ch in range(9):
            idx = np.where(p[ch] < thr)[0]
            want = SEL.longest_run(idx) if len(idx) else 0
            assert got[ch] == want
