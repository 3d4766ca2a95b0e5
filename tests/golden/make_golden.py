"""Generate the committed golden fixtures from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Writes ``tests/golden/*.npz``.  Inputs are stored as float32, reference outputs as
float64 exactly as the reference returned them (float32 where it preserves the
input dtype).  Versions of record are stored in each file's ``meta`` entry.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
from argparse import Namespace

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import reference_bridge as rb          # noqa: E402
from decode_tonal_langauge_b200 import synth       # noqa: E402
from decode_tonal_langauge_b200.chains import EX_STEPS, FULL6_STEPS  # noqa: E402


def meta():
    import scipy, pandas
    return json.dumps({"numpy": np.__version__, "scipy": scipy.__version__,
                       "pandas": pandas.__version__,
                       "reference": "Daniel-Lin-S/decode_tonal_langauge @ /root/reference"})


def ns(fs, **kw):
    return Namespace(signal_freq=fs, **kw)


def ref_chain(x, fs, steps):
    """The reference's own step modules called in sequence with a fresh parameter
    scope per step (its shared Namespace forbids repeated steps, Appendix B6)."""
    for step in steps:
        mod = rb.load("preprocess.signal." + step["module"].split(".")[-1])
        p = ns(fs, **(step.get("params") or {}))
        x = mod.run(x, p)
        fs = p.signal_freq
    return x, fs


def steps_fixture():
    fs, C, T = 2000, 3, 12000
    x, (on, tone, syl) = synth.session(0, C, T, fs, n_events=0)
    ff = rb.load("preprocess.signal.frequency_filter")
    out = {"x": x, "fs": fs}
    band = lambda m, **p: [{"method": m, "params": p}]
    out["notch"] = ff.run(x, ns(fs, bands=band("butter", freqs=[58, 62], filter_type="bandstop")))
    out["bandpass"] = ff.run(x, ns(fs, bands=band("butter", freqs=[70, 150], filter_type="bandpass")))
    out["lowpass"] = ff.run(x, ns(fs, bands=band("butter", freqs=200.0, filter_type="lowpass")))
    out["highpass"] = ff.run(x, ns(fs, bands=band("butter", freqs=1.0, filter_type="highpass")))
    out["causal"] = ff.run(x, ns(fs, bands=band("butter", freqs=[70, 150], filter_type="bandpass", causal=True)))
    out["hilbert_env"] = ff.run(x, ns(fs, bands=band("hilbert", freq_ranges=[70.0, 150.0], envelope=True)))
    out["hilbert_real"] = ff.run(x, ns(fs, bands=band("hilbert", freq_ranges=[70.0, 150.0], envelope=False)))
    out["hilbert_two_ranges"] = ff.run(x, ns(fs, bands=band(
        "hilbert", freq_ranges=[[30.0, 55.0], [70.0, 150.0]], envelope=True)))
    out["fir"] = ff.run(x, ns(fs, bands=band("fir", order=390, center_frequencies=[80.0, 100.0, 120.0])))
    out["two_bands"] = ff.run(x, ns(fs, bands=[
        {"method": "hilbert", "params": {"freq_ranges": [70.0, 150.0], "envelope": True}},
        {"method": "butter", "params": {"freqs": [0.3, 100], "filter_type": "bandpass"}}]))
    out["car"] = rb.load("preprocess.signal.car_rereference").run(x, ns(fs))
    out["car_excl"] = rb.load("preprocess.signal.car_rereference").run(x, ns(fs, exclude_channels=[1]))
    out["channel_zscore"] = rb.load("preprocess.signal.channel_zscore").run(x, ns(fs))
    out["zscore_rereference"] = rb.load("preprocess.signal.zscore_rereference").run(
        x, ns(fs, rereference_interval=[0.5, 3.0]))
    out["rolling_zscore"] = rb.load("preprocess.signal.rolling_zscore").run(x, ns(fs, window_length=1.5))
    p = ns(fs, downsample_freq=400)
    out["downsample"] = rb.load("preprocess.signal.downsample").run(x, p)
    out["downsample_fs"] = p.signal_freq
    # odd / non-5:1 lengths for the resampler's bin rules
    out["x_odd"] = x[:2, :9001]
    out["downsample_odd"] = rb.load("preprocess.signal.downsample").run(out["x_odd"], ns(fs, downsample_freq=400))
    out["downsample_600"] = rb.load("preprocess.signal.downsample").run(x[:2], ns(fs, downsample_freq=600))
    np.savez(os.path.join(HERE, "steps.npz"), meta=meta(), **out)
    print("steps.npz", {k: getattr(v, "shape", v) for k, v in out.items()})


def chains_fixture():
    fs, C, T = 2000, 3, 60000           # 30 s so EX's [0, 25] s interval exists
    x, _ = synth.session(1, C, T, fs)
    out = {"x": x, "fs": fs}
    y, f = ref_chain(x, fs, FULL6_STEPS)
    out["full6"], out["full6_fs"] = y, f
    y, f = ref_chain(x, fs, EX_STEPS)
    out["ex"], out["ex_fs"] = y, f
    np.savez(os.path.join(HERE, "chains.npz"), meta=meta(), **out)
    print("chains.npz", {k: getattr(v, "shape", v) for k, v in out.items()})


def epochs_fixture():
    import pandas as pd
    ta = rb.load("data_loading.text_align")
    rng = np.random.default_rng(7)
    C, sf_e, dur = 4, 400, 40.0
    sf_a = 2441.40625
    out = {}
    with tempfile.TemporaryDirectory() as d:
        intervals = {}
        for b in (1, 2):
            ecog = rng.standard_normal((C, int(sf_e * dur))).astype(np.float32)
            audio = rng.standard_normal((1, int(sf_a * dur))).astype(np.float32)
            np.savez(os.path.join(d, f"B{b}_ecog.npz"), data=ecog, sf=sf_e)
            np.savez(os.path.join(d, f"B{b}_sound.npz"), data=audio, sf=sf_a)
            n = 12
            starts = np.sort(rng.choice(np.arange(60, 380), n, replace=False)) / 10.0
            tone = rng.integers(1, 5, n)
            syl = rng.choice(["i", "a", "u"], n, p=[0.45, 0.45, 0.1])
            intervals[b] = pd.DataFrame({"start": starts, "end": starts + 0.6,
                                         "syllable": syl, "tone": tone})
            out[f"b{b}_ecog"], out[f"b{b}_audio"] = ecog, audio
            out[f"b{b}_start"], out[f"b{b}_tone"], out[f"b{b}_syllable"] = starts, tone, np.array(syl)
        res = ta.extract_ecog_audio(intervals, d, ["i", "a"], length=1.0,
                                    output_path=None, rest_period=(0.0, 5.0))
        # the reference merges blocks in os.listdir order; record it
        listing = [f for f in os.listdir(d)]
    for k, v in res.items():
        out["ref_" + k] = np.asarray(v)
    out["listdir"] = np.array(listing)
    out["ecog_sf"], out["audio_sf"] = sf_e, sf_a
    np.savez(os.path.join(HERE, "epochs.npz"), meta=meta(), **out)
    print("epochs.npz", {k: getattr(v, "shape", v) for k, v in out.items()})


def selection_fixture():
    disc = rb.load("channel_selection.discriminative")
    act = rb.load("channel_selection.active")
    rng = np.random.default_rng(11)
    N, C, L, R = 120, 10, 100, 9
    tone = rng.integers(0, 4, N).astype(np.int64)
    syl = rng.integers(0, 2, N).astype(np.int8)
    ecog = rng.standard_normal((N, C, L)) + 3.0
    win = np.zeros(L)
    win[20:80] = np.hanning(60)
    ecog[:, 0, :] += 1.2 * tone[:, None] * win          # tone-discriminative
    ecog[:, 1, :] += 0.5 * tone[:, None] * win
    ecog[:, 2, :] += 1.5 * syl[:, None] * win           # syllable-discriminative
    ecog[:, 3, :] += 2.0 * win                          # active, not discriminative
    ecog[:, 4, 40:52] += 2.0                            # short burst: run of 12
    ecog[:, 5, :] = 1.0                                 # constant everywhere -> NaN
    rest = rng.standard_normal((R, C, L)) + 3.0
    rest[:, 5, :] = 1.0
    # float32-representable values, so the float32 device path sees identical inputs
    ecog = ecog.astype(np.float32).astype(np.float64)
    rest = rest.astype(np.float32).astype(np.float64)
    data = {"ecog": ecog, "ecog_rest": rest, "ecog_sf": np.array(100), "tone": tone, "syllable": syl}
    out = dict(data)
    for target in ("tone", "syllable"):
        r = disc.run(data, {"p_threshold": 0.01, "active_time_threshold": 0.1, "target": target})
        out[f"disc_{target}_selected"] = np.array(r["selected_channels"], dtype=np.int64)
        out[f"disc_{target}_p"] = r["p_values"]
        out[f"disc_{target}_f"] = disc.test_discriminative_power(data, {"target": target})["f_stat"]
    r = act.run(data, {"p_threshold": 0.01, "active_time_threshold": 0.1})
    out["active_selected"] = np.array(r["selected_channels"], dtype=np.int64)
    out["active_max_lengths"] = np.array(r["max_lengths"], dtype=np.int64)
    out["active_p_last"] = r["p_values"]
    np.savez(os.path.join(HERE, "selection.npz"), meta=meta(), **out)
    print("selection.npz", {k: getattr(v, "shape", v) for k, v in out.items()})


if __name__ == "__main__":
    if not rb.available():
        raise SystemExit("needs the reference tree (build container only)")
    steps_fixture()
    chains_fixture()
    epochs_fixture()
    selection_fixture()
