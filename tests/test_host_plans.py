"""CPU tests of the host-built plans and tables through numpy models of the kernels
(tests/helpers/emulate.py).  No GPU and no compute calls into the library."""
import ctypes

import numpy as np
import pytest

from decode_tonal_langauge_b200 import design as D
from decode_tonal_langauge_b200 import fftplan as FP
from oracle import steps as S
from helpers import emulate as EM
from conftest import max_rel


@pytest.mark.parametrize("n", [1, 2, 5, 12, 60, 64, 225, 1000, 1800])
def test_tile_fft_model(n):
    rng = np.random.default_rng(n)
    x = rng.standard_normal((n, 3)) + 1j * rng.standard_normal((n, 3))
    got = EM.tile_fft(x, FP.axis_plan(n))
    ref = np.fft.fft(x, axis=0)
    assert np.max(np.abs(got - ref)) < 2e-6 * max(1.0, np.max(np.abs(ref)))


def test_factorize_rejects_non_smooth():
    with pytest.raises(NotImplementedError):
        FP.factorize(2 * 915527)
    assert np.prod(FP.factorize(3600000)) == 3600000


@pytest.mark.parametrize("N", [6000, 9000, 36000])
def test_big_fft_model(N):
    rng = np.random.default_rng(N)
    z = rng.standard_normal(N) + 1j * rng.standard_normal(N)
    plan = FP.big_plan(N)
    assert plan.a.n * plan.b.n == N and plan.a.n <= FP.MAX_AXIS and plan.b.n <= FP.MAX_AXIS
    got = EM.big_fft(z, plan)
    ref = np.fft.fft(z)
    assert np.max(np.abs(got - ref)) < 5e-6 * np.max(np.abs(ref))


def test_split_sizes_for_baseline_configs():
    for T in (1_200_000, 3_600_000, 7_200_000, 10_800_000):
        for N in (T // 2, T // 10 if T % 3 else T // 15):
            na, nb = FP.split_size(N)
            assert na * nb == N and max(na, nb) <= FP.MAX_AXIS


@pytest.mark.parametrize("T,num", [(12000, 2400), (12000, 3600), (16000, 16000), (6000, 9000), (24000, 4800)])
def test_resample_model_vs_oracle(T, num):
    rng = np.random.default_rng(T + num)
    x = rng.standard_normal(T).astype(np.float32) * 30
    got = EM.resample_model(x, FP.resample_plan(T, num))
    ref = S.fft_resample(x[None].astype(np.float64), num)[0]
    assert max_rel(got, ref) < 2e-6


@pytest.mark.parametrize("T,num", [(12000, 2400), (24000, 3200), (60000, 20000), (36000, 4800)])
def test_two_stage_resample_model_vs_oracle(T, num):
    """FIR pre-decimation + gain-compensated brick wall == scipy.signal.resample of the whole row."""
    pre = FP.predecimation(T, num)
    assert pre is not None and T % pre.D == 0 and pre.offset % 4 == 0 and len(pre.taps) % 4 == 0
    # ratios >= 5 with T % 4 == 0 take the two half-band stages (csrc/firdecim.cu: halfband2_decimate_kernel)
    assert (pre.halfband is not None) == (T % 4 == 0 and num / T <= 0.2), (T, num)
    if pre.halfband is not None:
        assert pre.D == 4 and len(pre.halfband[0]) - 1 <= FP.HALFBAND_K[0] and len(pre.halfband[1]) - 1 <= FP.HALFBAND_K[1]
    rng = np.random.default_rng(T + num)
    x = rng.standard_normal(T).astype(np.float32) * 30          # white: worst case for aliasing
    x1 = EM.fir_decimate_model(x, pre)
    got = EM.resample_model(x1, FP.resample_plan(T // pre.D, num), pre.bin_gain)
    ref = S.fft_resample(x[None].astype(np.float64), num)[0]
    assert max_rel(got, ref) < 3e-6


def test_predecimation_declines_small_ratios():
    assert FP.predecimation(24000, 12000) is None       # T/2 == num: no transition band
    assert FP.predecimation(16000, 16000) is None
    assert FP.predecimation(6000, 9000) is None         # up-sampling
    assert FP.predecimation(12001, 2400) is None        # odd T: no integer decimation of the circle
    assert FP.predecimation(12002, 2400).D == 2         # T/2 odd: the chirp-z stage takes the 6001-sample rows


def test_resample_odd_lengths_are_declared_unsupported():
    with pytest.raises(NotImplementedError):
        FP.resample_plan(9001, 1800)          # the smooth-length plan declines; the chirp-z plan takes over


@pytest.mark.parametrize("T,num", [(9001, 1800), (18310, 2399), (12000, 2400), (1000, 1501), (4099, 4099),
                                   (7919, 1000), (6000, 9001)])
def test_czt_resample_model_vs_oracle(T, num):
    """Bluestein plan (any length, odd, prime) == scipy.signal.resample of the whole row."""
    rng = np.random.default_rng(T + num)
    x = (np.cumsum(rng.standard_normal(T)) * 0.3 + rng.standard_normal(T) * 5).astype(np.float32)
    p = FP.czt_plan(T, num)
    assert p.M1 >= T + p.K - 1 and p.M2 >= p.K + num - 1 and p.M1 % 2 == 0 and p.M2 % 2 == 0
    got = EM.czt_resample_model(x, p)
    ref = S.fft_resample(x[None].astype(np.float64), num)[0]
    assert max_rel(got, ref) < 2e-6


def test_czt_plan_for_a_real_tdt_block():
    p = FP.czt_plan(1_831_054, 239_999)       # 600 s at 3051.7578125 Hz -> 400 Hz
    assert p.M1 == p.fft1.a.n * p.fft1.b.n and max(p.fft1.a.n, p.fft1.b.n) <= FP.MAX_AXIS_NARROW
    assert p.M2 == p.fft2.a.n * p.fft2.b.n


def test_hilbert_models(golden):
    from decode_tonal_langauge_b200 import _native as nat
    n = nat.lib.ecog_hilbert_twiddle_floats()
    tw = np.zeros(n, dtype=np.float32)
    assert nat.lib.ecog_hilbert_twiddles(tw.ctypes.data_as(ctypes.c_void_p)) == 0
    rng = np.random.default_rng(0)
    z = rng.standard_normal(4096) + 1j * rng.standard_normal(4096)
    got = EM.fft4096_model(z, tw)
    assert np.max(np.abs(got - np.fft.fft(z))) < 2e-5
    zs = np.zeros(4096, dtype=np.complex128)
    zs[:256] = z[:256]
    got = EM.fft4096_sparse_model(zs[:256], tw)             # fast path: 256 non-zero inputs
    assert np.max(np.abs(got - np.fft.fft(zs))) < 1e-5

    g = golden("steps")
    x, fs = g["x"], float(g["fs"])
    cfs, sds = D.gaussian_bank([70.0, 150.0])
    assert len(cfs) == 8 and abs(cfs[0] - 73.728) < 1e-3
    halo = FP.hilbert_halo(cfs, sds, fs, x.shape[1])
    plan = FP.hilbert_gain(cfs, sds, fs, True)
    assert plan[2] == 1 and plan[0].shape == (8, 256)          # high-gamma bank at 2 kHz: one row of bins
    y = np.stack([EM.hilbert_block_model(r, plan, halo, True) for r in x])
    assert max_rel(y, g["hilbert_env"]) < 1e-6
    plan = FP.hilbert_gain(cfs, sds, fs, False)
    assert (plan[1] == 0).all()
    y = np.stack([EM.hilbert_block_model(r, plan, halo, False) for r in x])
    assert max_rel(y, g["hilbert_real"]) < 1e-6
    assert FP.hilbert_gain(cfs, sds, 400.0, True)[2] == 4      # same bank at 400 Hz: wider in bins
    cfs2, sds2 = D.gaussian_bank([[30.0, 55.0], [70.0, 150.0]])
    halo2 = FP.hilbert_halo(cfs2, sds2, fs, x.shape[1])
    y = np.stack([EM.hilbert_block_model(r, FP.hilbert_gain(cfs2, sds2, fs, True), halo2, True) for r in x])
    assert max_rel(y, g["hilbert_two_ranges"]) < 1e-6
    with pytest.raises(NotImplementedError):
        c3, s3 = D.gaussian_bank([0.5, 4.0])
        FP.hilbert_halo(c3, s3, fs, x.shape[1])


@pytest.mark.parametrize("key,freqs,btype,tol", [
    ("notch", [58, 62], "bandstop", None),
    ("bandpass", [70, 150], "bandpass", 1e-6),
    ("lowpass", 200.0, "lowpass", 1e-6),
    ("highpass", 1.0, "highpass", 1e-6),
])
def test_sos_design_and_chunk_scan_model(golden, key, freqs, btype, tol):
    g = golden("steps")
    x, fs = g["x"], float(g["fs"])
    d = D.butter_design(freqs, fs, 4, False, btype)
    for chunk in (1024, 4096, 16384):
        M, tail = D.chunk_ops(d, chunk)
        y = EM.sos_chunk_scan(x, d, chunk, M, tail)
        if tol is not None:
            assert max_rel(y, g[key]) < tol
        else:   # ill-conditioned design: the long-double rule of SURVEY.md section 8c
            truth = S.filtfilt_pad(d.b, d.a, x, dtype=np.longdouble)
            assert max_rel(y, truth) <= max(1e-5, max_rel(g[key], truth))


@pytest.mark.parametrize("key,freqs,btype", [("notch", [58, 62], "bandstop"), ("bandpass", [70, 150], "bandpass")])
def test_sos_warm_up_model(golden, key, freqs, btype):
    """The warm-up path (zero-state start `tail` samples early, exact start-up near the row edge)
    meets the same bar as the exact carry scan."""
    g = golden("steps")
    x, fs = g["x"], float(g["fs"])
    d = D.butter_design(freqs, fs, 4, False, btype)
    tail = D.warm_tail(d, 1 << 20)
    assert tail > 0 and tail % 16 == 0
    for chunk in (1024, 2048):
        y = EM.sos_warm_model(x, d, chunk, tail)
        truth = S.filtfilt_pad(d.b, d.a, x, dtype=np.longdouble)
        assert max_rel(y, truth) <= max(1e-6, 1.05 * max_rel(g[key], truth))
    # long row: chunks far from the edges really start from zero
    rng = np.random.default_rng(5)
    xl = (rng.standard_normal((1, 60000)) * 30).astype(np.float32)
    if key == "bandpass":
        y = EM.sos_warm_model(xl, d, 4096, tail)
        assert max_rel(y, S.butter_filter(xl, freqs, fs, filter_type=btype)) < 1e-6


def test_warm_chunk_choice():
    L = D.choose_warm_chunk(256, 7_200_000, 512)
    assert L % 16 == 0 and 256 * -(-7_200_000 // L) <= 148 * 512
    d = D.butter_design([58, 62], 2000.0, 4, False, "bandstop")
    assert 0 < D.warm_tail(d, int(0.75 * L)) < 0.5 * L           # the C2 notch runs on the warm-up path
    assert D.warm_tail(d, 4096) == -1                            # short chunks: declined -> scan path


def test_causal_design_matches_reference(golden):
    g = golden("steps")
    x, fs = g["x"], float(g["fs"])
    d = D.butter_design([70, 150], fs, 4, True, "bandpass")
    M, tail = D.chunk_ops(d, 2048)
    y = EM.sos_chunk_scan(x, d, 2048, M, tail)
    assert max_rel(y, g["causal"]) < 1e-6


def test_choose_chunk_fills_one_wave():
    L = D.choose_chunk(256, 7_200_000)
    n_chunks = -(-7_200_000 // L)
    assert L % 16 == 0 and 256 * n_chunks <= 148 * 2 * 512
    assert D.choose_chunk(4, 12000) >= 1024


def test_numerator_form_follows_accuracy_not_coefficients():
    """Butterworth zeros sit on the unit circle, and b0 (1 + beta z^-1 + z^-2)^nsec equals the reference's
    rounded numerator to ~1e-16 per coefficient -- but next to the multiple zero the polynomial's value is
    ~1e-12, so that is NOT the same filter at a 4-Hz notch: the cheap unit form is used only where it gives
    the same output (design.NUMERATOR_TOL), otherwise the exact factors of the rounded (b, a)."""
    from scipy import signal as S
    unit = lambda sos: np.array_equal(sos[:, 2], sos[:, 0]) and np.all(sos[1:, 0] == 1.0)
    for freqs, fs, btype, want_unit in (([58, 62], 1000.0, "bandstop", True), ([58, 62], 2000.0, "bandstop", False),
                                        ([58, 62], 3051.7578125, "bandstop", False),
                                        (200.0, 2000.0, "lowpass", True), (1.0, 400.0, "highpass", True)):
        d = D.butter_design(freqs, fs, 4, False, btype)
        sos = d.sos
        assert unit(sos) == want_unit, (freqs, fs, btype)
        if want_unit:
            assert np.all(sos[:, 1] / sos[:, 0] == sos[1, 1])                 # one beta for every section
        b, _ = S.butter(4, np.asarray(freqs, dtype=float) / (fs / 2), btype=btype)
        prod = np.array([1.0])
        for sec in sos:
            prod = np.convolve(prod, sec[:3])
        assert np.max(np.abs(prod - b)) <= 4e-15 * np.max(np.abs(b))          # either way: the rounded b to rounding
    d = D.butter_design([70, 150], 2000.0, 4, False, "bandpass")
    assert np.array_equal(d.sos[:, 2], -d.sos[:, 0]) and np.all(d.sos[:, 1] == 0.0)


def test_notch_3khz_cascade_matches_extended_precision_reference():
    """The worst-conditioned design of SURVEY C2 (58-62 Hz at 3 kHz, pole radius 0.9985): the device
    cascade (numpy model, float64 state, float32 storage) against the reference's own algorithm evaluated
    in long double with the reference's float64 lfilter_zi -- 2e-7, while the float64 reference itself is
    3e-4 away from it."""
    from oracle import steps as OS
    from scipy import signal as S
    fs, T = 3000.0, 90_000
    rng = np.random.default_rng(0)
    t = np.arange(T) / fs
    x = (rng.standard_normal((2, T)) * 30 + 10 * np.sin(2 * np.pi * 60 * t)).astype(np.float32)
    d = D.butter_design([58, 62], fs, 4, False, "bandstop")
    got = EM.sos_warm_model(x, d, T, 0)
    zi = OS.lfilter_zi(d.b, d.a).astype(np.longdouble)
    b, a = np.asarray(d.b, dtype=np.longdouble), np.asarray(d.a, dtype=np.longdouble)
    ext = OS.odd_extension(x, 27).astype(np.longdouble)
    y, _ = S.lfilter(b, a, ext, axis=-1, zi=zi[None] * ext[..., :1])
    y, _ = S.lfilter(b, a, y[..., ::-1], axis=-1, zi=zi[None] * y[..., -1:])
    truth = np.asarray(y[..., ::-1][..., 27:-27], dtype=np.float64)
    ref = OS.filtfilt_pad(d.b, d.a, x)
    assert max_rel(got, truth) < 1e-6 < 1e-4 < max_rel(ref, truth)


def test_pair_design_forms():
    """The fused cascade pair: forms (exact notch, unit band-pass) -> band-pass monic, both gains on the
    notch's first section; its unit-step state is scipy's sosfilt_zi of that cascade."""
    from scipy import signal as S
    A = D.butter_design([58, 62], 2000.0, 4, False, "bandstop")
    B = D.butter_design([70, 150], 2000.0, 4, False, "bandpass")
    dsg, tail_b = D.pair_design(A, B)
    assert dsg.sos.shape == (8, 6) and np.all(dsg.sos[4:, :3] == [1.0, 0.0, -1.0])
    assert np.isclose(dsg.sos[0, 0], A.sos[0, 0] * B.sos[0, 0], rtol=1e-15)
    assert np.allclose(dsg.zi, S.sosfilt_zi(dsg.sos)) and tail_b == 2 * D.warm_tail(B, 1 << 30)
    u = np.random.default_rng(3).standard_normal(5000)
    assert np.allclose(S.sosfilt(dsg.sos, u), S.sosfilt(B.sos, S.sosfilt(A.sos, u)), rtol=1e-9, atol=1e-12)
    lp = D.butter_design(100.0, 2000.0, 4, False, "lowpass")
    assert D.pair_design(A, lp) is None                                      # 2 sections: no pair kernel


def test_hilbert_nz_marks_the_zero_tail():
    cfs, sds = D.gaussian_bank([70.0, 150.0])
    for fs in (2000.0, 3000.0):
        gain, shift, rows = FP.hilbert_gain(cfs, sds, fs, True)
        assert rows == 1
        nz = FP.hilbert_nz(gain)
        assert nz.shape == (8,) and nz.min() >= 1 and nz.max() <= 16
        for g, n in zip(gain, nz):
            assert not np.any(g[16 * n:]) and np.any(g[16 * (n - 1):16 * n])
        # the dropped tails are below 1e-9 of the band peak
        f = np.arange(2048) * fs / 4096
        for b in range(8):
            full = np.exp(-0.5 * ((f - cfs[b]) / sds[b]) ** 2)
            kept = np.zeros(2048, dtype=bool)
            kept[shift[b]:shift[b] + 256] = gain[b] != 0
            assert full[~kept].max() < 1e-9 * full.max()


def test_bind_host_to_gpu_is_best_effort():
    """No GPU / no NVML here: the affinity helper must decline quietly and leave the mask alone."""
    import os
    from decode_tonal_langauge_b200 import runtime as rt
    before = os.sched_getaffinity(0)
    got = rt.bind_host_to_gpu(0)
    assert got is None or set(got) <= before
    assert os.sched_getaffinity(0) == (set(got) if got else before)


def test_bandpass_delta_form_is_the_same_filter():
    """The float32 half of the cascade pair (csrc/sos_common.cuh: Bp32) carries w[n-1] and w[n-1] - w[n-2]
    with the coefficients c1 = -(1 + a1 + a2), e2 = 1 - a2.  In float64 the recursion must reproduce
    scipy's sosfilt of the same band-pass; in float32 it must stay within the bound design.bandpass_f32_ok
    promises (<= ~1e-6 of the row maximum)."""
    from scipy import signal as S
    from decode_tonal_langauge_b200 import design as D
    fs = 2000.0
    B = D.butter_design([70, 150], fs, 4, False, "bandpass")
    assert D.bandpass_f32_ok(B)
    assert not D.bandpass_f32_ok(D.butter_design([58, 62], fs, 4, False, "bandstop"))
    assert not D.bandpass_f32_ok(D.butter_design([70, 150], 6000.0, 4, False, "bandpass"))
    sos = np.asarray(B.sos, dtype=np.float64)
    x = np.random.default_rng(3).standard_normal(6000) * 30.0
    ref = S.sosfilt(sos, x)

    def delta(dtype):
        c1 = (-(1.0 + sos[:, 4] + sos[:, 5])).astype(dtype)
        e2 = (1.0 - sos[:, 5]).astype(dtype)
        g = dtype(np.prod(sos[:, 0]))
        w1 = np.zeros(4, dtype=dtype)
        d = np.zeros(4, dtype=dtype)
        y = np.empty(x.size, dtype=dtype)
        for n in range(x.size):
            v = dtype(x[n]) * g
            for j in range(4):
                dn = dtype(c1[j] * w1[j] + (v - e2[j] * d[j])) + d[j]
                v = dn + d[j]
                w1[j] = w1[j] + dn
                d[j] = dn
            y[n] = v
        return y.astype(np.float64)

    scale = np.max(np.abs(ref))
    assert np.max(np.abs(delta(np.float64) - ref)) / scale < 1e-12
    assert np.max(np.abs(delta(np.float32) - ref)) / scale < 1.5e-6


@pytest.mark.parametrize("T,num", [(7_200_000, 1_440_000), (10_800_000, 1_440_000), (1_200_000, 120_000)])
def test_halfband_predecimation_design(T, num):
    """fftplan.predecimation for ratios >= 5: two half-band Kaiser stages that fit the kernel (csrc/firdecim.cu:
    kHbK1 = 8, kHbK2 = 21), >= 120 dB where the decimation folds onto the kept band, and a compensation table that is
    the reciprocal of the product of the two (float32) stage responses."""
    pre = FP.predecimation(T, num)
    assert pre is not None and pre.D == 4 and pre.halfband is not None
    s1, s2 = pre.halfband
    assert s1.dtype == np.float32 and s2.dtype == np.float32
    assert len(s1) - 1 <= FP.HALFBAND_K[0] and len(s2) - 1 <= FP.HALFBAND_K[1]
    f_pass = num / T

    def resp(st, f):            # f in units of the stage's own Nyquist frequency
        i = np.arange(len(st) - 1)
        return float(st[0]) + 2 * np.sum(st[1:].astype(np.float64)[None] * np.cos(np.pi * f[:, None] * (2 * i + 1)[None]), axis=1)

    for st, fp in ((s1, f_pass), (s2, 2 * f_pass)):
        stop = np.abs(resp(st, np.linspace(1 - fp, 1, 2001)))
        assert 20 * np.log10(stop.max()) <= -120.0
        assert abs(resp(st, np.zeros(1))[0] - 1.0) < 1e-6          # DC gain 1
    k = np.arange(0, num // 2 + 1, max(1, num // 2000), dtype=np.float64)
    H = resp(s1, 2 * k / T) * resp(s2, 4 * k / T)
    assert np.allclose(pre.bin_gain[k.astype(int)], 1.0 / H, rtol=2e-6)


def test_pair_ws_shape_fills_the_sms():
    """ops.pair_ws_shape: chunks per CTA and chunk length of the warp-specialised cascade pair -- the chunk length
    divides the row, is a multiple of 32 and covers the warm-up; at C2 224 chunks per CTA give 143 CTAs for 148 SMs."""
    ops = pytest.importorskip("decode_tonal_langauge_b200.ops")
    for C, T, tail in ((256, 7_200_000, 10368), (128, 10_800_000, 16224), (21, 7_200_000, 10368), (8, 600_000, 10368)):
        shape = ops.pair_ws_shape(C, T, tail)
        assert shape is not None
        P, L = shape
        assert P in ops.PAIR_WS_CHUNKS and T % L == 0 and L % 32 == 0 and L >= tail
    assert ops.pair_ws_shape(256, 7_200_000, 10368) == (224, 57600)
    assert ops.pair_ws_shape(4, 7_200_002, 10368) is None          # no divisor that is a multiple of 32


def test_pair_ws_model_matches_two_filtfilts():
    """The algorithm of csrc/sosfilt_pairws.cu on the CPU (tests/helpers/emulate.py::pair_ws_model): one forward and
    one backward sweep of notch (float64) -> band-pass (float32 delta form), zero-state warm-up per chunk, against
    scipy's filtfilt(band-pass, filtfilt(notch, x)) away from the row ends (the product overwrites the ends)."""
    from decode_tonal_langauge_b200 import design as D
    from decode_tonal_langauge_b200 import synth
    fs, T, L = 2000.0, 48_000, 12_000
    A = D.butter_design([58, 62], fs, 4, False, "bandstop")
    B = D.butter_design([70, 150], fs, 4, False, "bandpass")
    comb = D.pair_design(A, B)
    assert comb is not None and D.bandpass_f32_ok(B)
    natural = D.SosDesign(np.ascontiguousarray(np.vstack([A.sos, B.sos])), None, comb[0].padlen, True)
    tail = D.warm_tail(natural, T)
    assert 0 < tail <= L
    tail32 = -(-tail // 32) * 32
    x = synth.session_channels(1, [5], T, fs, 256)[0]
    got = EM.pair_ws_model(x, A, B, L, tail32, min(comb[1], tail32))
    # the notch of the truth in long double: the float64 direct form of the reference is itself 8e-6 off at 2 kHz
    n_ld = np.asarray(S.filtfilt_pad(A.b, A.a, x[None].astype(np.float64), dtype=np.longdouble, reference_edges=True), dtype=np.float64)
    ref = S.filtfilt_pad(B.b, B.a, n_ld)[0]
    inner = slice(tail32, T - tail32)
    err = np.max(np.abs(got[inner] - ref[inner])) / np.max(np.abs(ref[inner]))
    assert err < 2e-6, err
