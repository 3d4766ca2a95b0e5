"""Oracle vs the LIVE reference imported from /root/reference (build container only;
skipped on the GPU box where the tree does not exist)."""
from argparse import Namespace

import numpy as np
import pytest

from oracle import reference_bridge as rb
from oracle import steps as S
from oracle import selection as SEL
from decode_tonal_langauge_b200 import synth
from conftest import max_rel

pytestmark = pytest.mark.skipif(not rb.available(), reason="reference tree not mounted")


@pytest.fixture(scope="module")
def x():
    return synth.session(3, 5, 9000, 3000.0)[0]


def test_steps_live(x):
    fs = 3000.0
    ff = rb.load("preprocess.signal.frequency_filter")
    for bands in ([{"method": "butter", "params": {"freqs": [58, 62], "filter_type": "bandstop"}}],
                  [{"method": "butter", "params": {"freqs": [70, 150], "filter_type": "bandpass"}}],
                  [{"method": "hilbert", "params": {"freq_ranges": [70.0, 150.0]}}]):
        ref = ff.run(x, Namespace(signal_freq=fs, bands=bands))
        assert max_rel(S.frequency_filter(x, fs, bands), ref) < 1e-11
    p = Namespace(signal_freq=fs, downsample_freq=400)
    ref = rb.load("preprocess.signal.downsample").run(x, p)
    y, f = S.downsample(x, fs, 400)
    assert f == p.signal_freq and np.array_equal(y, ref)
    assert np.array_equal(S.car_rereference(x, [0, 2]),
                          rb.load("preprocess.signal.car_rereference").run(
                              x, Namespace(signal_freq=fs, exclude_channels=[0, 2])))


def test_hilbert_int_range_is_a_reference_defect(x):
    """Appendix B2: the shipped example's ``freq_ranges: [70, 150]`` raises in the reference."""
    ff = rb.load("preprocess.signal.frequency_filter")
    with pytest.raises(TypeError):
        ff.run(x, Namespace(signal_freq=3000.0, bands=[{"method": "hilbert",
                                                         "params": {"freq_ranges": [70, 150]}}]))


def test_selection_live():
    rng = np.random.default_rng(5)
    N, C, L = 90, 6, 50
    data = {"ecog": rng.standard_normal((N, C, L)), "ecog_rest": rng.standard_normal((7, C, L)),
            "ecog_sf": np.array(100), "tone": rng.integers(0, 3, N)}
    data["ecog"][:, 1, 10:40] += data["tone"][:, None] * 1.0
    params = {"p_threshold": 0.01, "active_time_threshold": 0.1, "target": "tone"}
    ref = rb.load("channel_selection.discriminative").run(data, params)
    mine = SEL.discriminative(data, params)
    assert mine["selected_channels"] == ref["selected_channels"]
    assert np.allclose(mine["p_values"], ref["p_values"], rtol=1e-9, atol=0)
    ref = rb.load("channel_selection.active").run(data, params)
    mine = SEL.active(data, params)
    assert mine["selected_channels"] == ref["selected_channels"]
    assert mine["max_lengths"] == ref["max_lengths"]
