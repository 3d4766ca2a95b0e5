"""The oracle restatement against the committed fixtures written by the REAL reference
(tests/golden/make_golden.py).  CPU only; runs everywhere."""
import numpy as np
import pytest

from oracle import steps as S
from oracle import epochs as E
from oracle import selection as SEL
from oracle import chains as CH
from decode_tonal_langauge_b200.chains import EX_STEPS, FULL6_STEPS
from conftest import max_rel

TIGHT = 1e-11          # float64 restatement of float64 reference arithmetic


def band(m, **p):
    return [{"method": m, "params": p}]


@pytest.fixture(scope="module")
def g(golden):
    return golden("steps")


@pytest.mark.parametrize("key,bands", [
    ("notch", band("butter", freqs=[58, 62], filter_type="bandstop")),
    ("bandpass", band("butter", freqs=[70, 150], filter_type="bandpass")),
    ("lowpass", band("butter", freqs=200.0, filter_type="lowpass")),
    ("highpass", band("butter", freqs=1.0, filter_type="highpass")),
    ("causal", band("butter", freqs=[70, 150], filter_type="bandpass", causal=True)),
    ("hilbert_env", band("hilbert", freq_ranges=[70.0, 150.0], envelope=True)),
    ("hilbert_real", band("hilbert", freq_ranges=[70.0, 150.0], envelope=False)),
    ("hilbert_two_ranges", band("hilbert", freq_ranges=[[30.0, 55.0], [70.0, 150.0]], envelope=True)),
    ("two_bands", [{"method": "hilbert", "params": {"freq_ranges": [70.0, 150.0], "envelope": True}},
                   {"method": "butter", "params": {"freqs": [0.3, 100], "filter_type": "bandpass"}}]),
])
def test_frequency_filter(g, key, bands):
    y = S.frequency_filter(g["x"], int(g["fs"]), bands)
    assert y.dtype == g[key].dtype
    assert max_rel(y, g[key]) < TIGHT


def test_fir(g):
    y = S.frequency_filter(g["x"], int(g["fs"]),
                           band("fir", order=390, center_frequencies=[80.0, 100.0, 120.0]))
    assert y.dtype == g["fir"].dtype == np.float32
    assert max_rel(y, g["fir"]) < 1e-6


def test_car_and_zscores(g):
    x, fs = g["x"], int(g["fs"])
    assert np.array_equal(S.car_rereference(x), g["car"])
    assert np.array_equal(S.car_rereference(x, [1]), g["car_excl"])
    assert np.array_equal(S.channel_zscore(x), g["channel_zscore"])
    assert np.array_equal(S.zscore_rereference(x, fs, [0.5, 3.0]), g["zscore_rereference"])
    r = S.rolling_zscore(x, fs, 1.5)
    assert np.isnan(r[:, 0]).all() and np.isnan(g["rolling_zscore"][:, 0]).all()
    assert max_rel(r[:, 1:], g["rolling_zscore"][:, 1:]) < 1e-9


def test_downsample(g):
    x, fs = g["x"], int(g["fs"])
    y, f = S.downsample(x, fs, 400)
    assert f == int(g["downsample_fs"]) and y.dtype == np.float32
    assert np.array_equal(y, g["downsample"])
    assert np.array_equal(S.downsample(g["x_odd"], fs, 400)[0], g["downsample_odd"])
    assert np.array_equal(S.downsample(x[:2], fs, 600)[0], g["downsample_600"])


def test_errors():
    x = np.zeros((2, 20))
    with pytest.raises(ValueError):
        S.filtfilt_pad([1, 2, 1], [1, -0.5, 0.1], x[:, :9])
    with pytest.raises(ValueError):
        S.car_rereference(x, [5])
    with pytest.raises(ValueError):
        S.zscore_rereference(x, 10, [1.0, 3.0])
    with pytest.raises(ValueError):
        S.zscore_rereference(x, 10, [1.0, 1.0])
    with pytest.raises(ValueError):
        S.rolling_zscore(x, 10, 0.1)


def test_chains(golden):
    c = golden("chains")
    y, f = CH.run_chain(c["x"], int(c["fs"]), FULL6_STEPS)
    assert f == int(c["full6_fs"]) and max_rel(y, c["full6"]) < 1e-9
    y, f = CH.run_chain(c["x"], int(c["fs"]), EX_STEPS)
    assert f == int(c["ex_fs"]) and max_rel(y, c["ex"]) < 1e-9


def test_epochs(golden):
    e = golden("epochs")
    intervals, rec = {}, {}
    for b in (1, 2):
        intervals[b] = {"start": e[f"b{b}_start"], "tone": e[f"b{b}_tone"],
                        "syllable": [str(s) for s in e[f"b{b}_syllable"]]}
        rec[b] = {"ecog": (e[f"b{b}_ecog"], e["ecog_sf"][()]), "audio": (e[f"b{b}_audio"], e["audio_sf"][()])}
    order = []
    for f in e["listdir"]:
        if "sound" in str(f):
            order.append(int(str(f)[1]))
    out = E.extract_epochs(intervals, rec, ["i", "a"], 1.0, (0.0, 5.0))
    if order != sorted(order):      # reference merged in os.listdir order; re-merge the same way
        n = len(e["b1_start"])
        perm = np.concatenate([np.arange(n) + n * (b - 1) for b in order])
    else:
        perm = np.arange(out["ecog"].shape[0])
    assert np.array_equal(out["ecog"][perm], e["ref_ecog"])
    assert np.array_equal(out["audio"][perm], e["ref_audio"])
    assert np.array_equal(out["syllable"][perm], e["ref_syllable"]) and out["syllable"].dtype == np.int8
    assert np.array_equal(out["tone"][perm], e["ref_tone"])
    assert out["ecog_rest"].shape == e["ref_ecog_rest"].shape
    assert np.array_equal(np.sort(out["ecog_rest"].ravel()), np.sort(e["ref_ecog_rest"].ravel()))


def test_epoch_index_truncation():
    # int(2.3*400) == 919, int(4.1*400) == 1639
    first, n = E.onset_indices([2.3, 4.1, 1.0], 400, 1.0)
    assert first.tolist() == [919, 1639, 400] and n == 400
    first, n = E.onset_indices([2.3], 24414.0625, 1.0)
    assert first.tolist() == [int(2.3 * 24414.0625)] and n == 24414
    with pytest.raises(ValueError):
        E.gather(np.zeros((2, 100)), np.array([50]), 60)


def test_selection(golden):
    s = golden("selection")
    data = {k: s[k] for k in ("ecog", "ecog_rest", "ecog_sf", "tone", "syllable")}
    for target in ("tone", "syllable"):
        r = SEL.discriminative(data, {"p_threshold": 0.01, "active_time_threshold": 0.1, "target": target})
        assert r["selected_channels"] == s[f"disc_{target}_selected"].tolist()
        assert max_rel(r["f_stat"], s[f"disc_{target}_f"]) < 1e-10
        p, pr = r["p_values"], s[f"disc_{target}_p"]
        assert np.array_equal(np.isnan(p), np.isnan(pr))
        ok = ~np.isnan(pr)
        assert np.allclose(p[ok], pr[ok], rtol=1e-9, atol=0)
    r = SEL.active(data, {"p_threshold": 0.01, "active_time_threshold": 0.1})
    assert r["selected_channels"] == s["active_selected"].tolist()
    assert r["max_lengths"] == s["active_max_lengths"].tolist()
    assert np.allclose(r["p_values"], s["active_p_last"], rtol=1e-9, atol=0)
