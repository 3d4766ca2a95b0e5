"""Host logic of the stage drivers and plug-in host: naming, provenance, TextGrid parsing,
parameter scoping.  CPU only."""
import os
from argparse import Namespace

import numpy as np
import pytest
import yaml

from decode_tonal_langauge_b200 import config as cfgmod
from decode_tonal_langauge_b200 import stages, textgrid_io
from decode_tonal_langauge_b200.preprocessor import preprocess_signal

EXAMPLE = {
    "preprocess": {"module": "preprocess_main", "params": {"modalities": {"ecog": {"type": "signal", "preprocessing": {
        "steps": [
            {"module": "preprocess.downsample", "params": {"downsample_freq": 400}},
            {"module": "preprocess.frequency_filter", "params": {"bands": [
                {"method": "hilbert", "params": {"freq_ranges": [70, 150], "envelope": True}},
                {"method": "butter", "params": {"freqs": [0.3, 100], "filter_type": "bandpass"}}]}},
            {"module": "preprocess.zscore_rereference", "params": {"rereference_interval": [0.0, 25.0]}}]}},
        "audio": {"type": "signal"}}}},
}


def test_setup_name_matches_reference_for_example_config():
    # value produced by the reference's generate_setup_name on example_config.yaml:27-45
    assert stages.generate_setup_name(EXAMPLE["preprocess"]["params"]["modalities"]) == \
        "downsample__frequency_filter__zscore_rereference_4d5793"
    assert stages.generate_setup_name({"audio": {"type": "signal"}}) == "raw"


def test_hash_names_against_live_reference():
    from oracle import reference_bridge as rb
    if not rb.available():
        pytest.skip("reference tree not mounted")
    cfg = yaml.safe_load(open(os.path.join(rb.REFERENCE_ROOT, "example_config.yaml")))
    sb = rb.load("preprocess.pipelines.subject_block")
    assert stages.generate_setup_name(cfg["preprocess"]["params"]["modalities"]) == \
        sb.generate_setup_name(cfg["preprocess"]["params"]["modalities"])
    uc = rb.load("utils.config")
    assert cfgmod.generate_hash_name_from_config("s__1", cfg["channel_selection"]) == \
        uc.generate_hash_name_from_config("s__1", cfg["channel_selection"])
    es = rb.load("extract_samples")
    assert cfgmod.yaml_hash_name("setup", cfg["sample_collection"]) == \
        es._generate_output_dir_name("setup", cfg["sample_collection"])
    ns = cfgmod.dict_to_namespace({"a": {"b": [1, {"c": 2}]}, "root_dir": {"x": 1}}, exclude_keys=["root_dir"])
    ref = uc.dict_to_namespace({"a": {"b": [1, {"c": 2}]}, "root_dir": {"x": 1}}, exclude_keys=["root_dir"])
    assert ns == ref and ns.root_dir == {"x": 1}


def test_update_configuration_merges_provenance(tmp_path):
    prev = tmp_path / "prev.yaml"
    prev.write_text(yaml.dump({"preprocess": {"io": {"root_dir": "r"}}}))
    out = tmp_path / "config.yaml"
    cfgmod.update_configuration(str(out), str(prev), "sample_collection", {"params": {"x": 1}})
    merged = yaml.safe_load(out.read_text())
    assert merged == {"preprocess": {"io": {"root_dir": "r"}}, "sample_collection": {"params": {"x": 1}}}
    cfgmod.update_configuration(str(out), str(tmp_path / "missing.yaml"), "channel_selection", {})
    assert yaml.safe_load(out.read_text()) == {"channel_selection": {}}


LONG_TG = '''File type = "ooTextFile"
Object class = "TextGrid"

xmin = 0
xmax = 12.5
tiers? <exists>
size = 2
item []:
    item [1]:
        class = "IntervalTier"
        name = "Success"
        xmin = 0
        xmax = 12.5
        intervals: size = 5
        intervals [1]:
            xmin = 0
            xmax = 1.234
            text = ""
        intervals [2]:
            xmin = 1.234
            xmax = 1.9
            text = "3i"
        intervals [3]:
            xmin = 1.9
            xmax = 4.06
            text = "rest"
        intervals [4]:
            xmin = 4.06
            xmax = 4.71
            text = "1a said ""ma"""
        intervals [5]:
            xmin = 4.5
            xmax = 5.0
            text = "2i"
    item [2]:
        class = "IntervalTier"
        name = "fail"
        xmin = 0
        xmax = 12.5
        intervals: size = 1
        intervals [1]:
            xmin = 7.0
            xmax = 7.5
            text = "4a"
'''

SHORT_TG = '''File type = "ooTextFile"
Object class = "TextGrid"

0
12.5
<exists>
1
"IntervalTier"
"success"
0
12.5
2
0.52
1.0
"2a"
1.0
12.5
""
'''


def test_textgrid_parser_and_interval_table(tmp_path):
    tg = textgrid_io.TextGrid.fromString(LONG_TG)
    assert [t.name for t in tg.tiers] == ["Success", "fail"]
    assert len(tg.tiers[0].intervals) == 5 and tg.tiers[0].intervals[3].mark == '1a said "ma"'
    df = textgrid_io.read_textgrid(tg, start_offset=0.2, end_offset=0.0, tier_list=["success"])
    # third trial (4.5 s) overlaps the previous one's end (4.7) after the offset -> skipped
    assert df["start"].tolist() == [1.0, 3.9] and df["end"].tolist() == [1.9, 4.7]
    assert df["tone"].tolist() == [3, 1] and df["syllable"].tolist() == ["i", "a"]
    assert df["start"].dtype == np.float64
    all_tiers = textgrid_io.read_textgrid(tg, 0.0, 0.0, None)
    assert len(all_tiers) == 3 and all_tiers["tone"].tolist()[-1] == 4
    short = textgrid_io.TextGrid.fromString(SHORT_TG)
    assert short.tiers[0].intervals[0].mark == "2a" and short.tiers[0].intervals[0].minTime == 0.52
    (tmp_path / "x_B2.TextGrid").write_text(LONG_TG)
    (tmp_path / "y_B7.TextGrid").write_text(SHORT_TG)
    (tmp_path / "notes.txt").write_text("ignored")
    tables = textgrid_io.handle_textgrids(str(tmp_path), start_offset=0.2, tier_list=["success"], blocks=[2, 7])
    assert sorted(tables) == [2, 7] and len(tables[7]) == 1 and tables[7]["start"].iloc[0] == 0.3
    assert list(textgrid_io.handle_textgrids(str(tmp_path), blocks=[7])) == [7]


def test_step_runner_param_scope_and_rate_threading():
    x = np.ones((2, 8), dtype=np.float32)
    steps = [{"module": "helpers.fake_step", "params": {"gain": 2.0}},
             {"module": "helpers.fake_step", "params": {"gain": 3.0, "halve_rate": True}}]
    p = Namespace(signal_freq=100)
    y, f = preprocess_signal(x, steps, p)              # repeated step + repeated key: allowed (B6)
    assert f == 50 and y.shape == (2, 4) and np.all(y == 6.0) and p.signal_freq == 50
    # the reference flattens every step's params into ONE shared Namespace (preprocessor.py:46-53):
    # keys of earlier steps stay visible to later ones and to the caller; a repeated key is overwritten
    assert p.gain == 3.0 and p.halve_rate is True
    # ... so a key set by step 1 only governs step 2 as well (the reference's behaviour)
    steps2 = [{"module": "helpers.fake_step", "params": {"gain": 2.0, "halve_rate": True}},
              {"module": "helpers.fake_step", "params": {"gain": 3.0}}]
    y2, f2 = preprocess_signal(x, steps2, Namespace(signal_freq=100))
    assert f2 == 25 and y2.shape == (2, 2)
    with pytest.raises(ValueError, match="already exists"):
        preprocess_signal(x, steps, Namespace(signal_freq=100), strict_params=True)


def test_iter_blocks_and_block_ids(tmp_path):
    for d in ("HS1-B2", "HS1-B10", "HS1-notes", "HS1-3"):
        (tmp_path / "Sub1" / d).mkdir(parents=True)
    got = list(stages.iter_blocks(str(tmp_path), ["Sub1"], [7]))
    assert [(s, b) for s, b, _ in got] == [(7, 3), (7, 10), (7, 2)] or sorted(b for _, b, _ in got) == [2, 3, 10]
    assert stages.get_block_id("HS1-B12") == 12 and stages.get_block_id("HS1-x") is None


def test_stage_io_wiring():
    outs = {"preprocess": "/p", "sample_collection": "/s", "channel_selection": "/c"}
    sc = {}
    stages._wire_io(outs, "sample_collection", sc)
    assert sc["params"]["io"]["recording_dir"] == "/p"
    sc = {"params": {"io": {"sample_dir": "/mine"}}}
    stages._wire_io(outs, "channel_selection", sc)
    assert sc["params"]["io"]["sample_dir"] == "/mine"
    sc = {}
    stages._wire_io(outs, "training", sc)
    assert sc["params"]["io"] == {"sample_dir": "/s", "channel_selection_dir": "/c"}


def test_dropin_names_resolve():
    import importlib
    import decode_tonal_langauge_b200 as pkg
    pkg.install_dropin()
    for name in ("preprocess.downsample", "preprocess.signal.downsample", "preprocess.frequency_filter",
                 "preprocess.zscore_rereference", "preprocess.car_rereference", "preprocess.channel_zscore",
                 "preprocess.rolling_zscore", "preprocess.preprocessor", "preprocess.pipelines.subject_block",
                 "preprocess.io.tdt_blocks", "preprocess_main", "extract_samples", "channel_selection_main",
                 "channel_selection.active", "channel_selection.discriminative", "data_loading.text_align", "main"):
        mod = importlib.import_module(name)
        assert mod.__file__.startswith(pkg.DROPIN_DIR), (name, mod.__file__)
    assert callable(importlib.import_module("preprocess.downsample").run)
    assert callable(importlib.import_module("preprocess_main").run)         # Appendix B4
    from channel_selection.utils import get_max_length
    assert get_max_length(np.array([1, 2, 3, 7, 8, 10, 11, 12, 13])) == 4


def test_sample_handler_host_logic():
    """Joint label code and channel union of the downstream consumer (contract of
    ref: data_loading/sample_loading.py:66-72,101-121), no device involved."""
    from decode_tonal_langauge_b200.samples import joint_codes, selected_union
    tone = np.array([0, 3, 1, 2, 3])
    syl = np.array([1, 0, 0, 1, 1], dtype=np.int8)
    assert joint_codes([tone, syl]).tolist() == (tone + 4 * syl).tolist()
    assert joint_codes([syl, tone]).tolist() == (syl + 2 * tone).tolist()
    assert joint_codes([tone]).dtype == int
    sel = {"tone_discriminative": [9, 3], "syllable_discriminative": [3, 1], "active": [0]}
    assert selected_union(sel, ["tone", "syllable"]).tolist() == [1, 3, 9]
    with pytest.raises(KeyError, match="nope_discriminative"):
        selected_union(sel, ["nope"])
    with pytest.raises(ValueError, match="No channels"):
        selected_union({"tone_discriminative": []}, ["tone"])


def test_npz_member_is_read_in_place(tmp_path):
    """sessions.locate_npz_array / _read_into: the raw bytes of `data` inside the reference's block format
    (np.savez: stored zip member + .npy header) go straight into a caller-owned buffer."""
    from decode_tonal_langauge_b200 import sessions as SS
    rng = np.random.default_rng(0)
    for dt, shape in ((np.float32, (5, 70001)), (np.float64, (3, 1000)), (np.int16, (2, 33))):
        data = (rng.standard_normal(shape) * 100).astype(dt)
        path = tmp_path / f"B1_ecog_{np.dtype(dt).name}.npz"
        np.savez(path, data=data, sf=np.float64(2000.0))
        m = SS.locate_npz_array(str(path), "data")
        assert m.shape == shape and m.dtype == np.dtype(dt) and not m.fortran
        out = np.empty(shape, dtype=dt)
        SS._read_into(m, out, threads=3)
        assert np.array_equal(out, data)
        assert SS.read_npz_scalar(str(path), "sf") == 2000.0
    big = rng.standard_normal((4, 5_000_000)).astype(np.float32)          # > 64 MiB: the threaded path
    np.savez(tmp_path / "big.npz", data=big, sf=np.int64(400))
    out = np.empty_like(big)
    SS._read_into(SS.locate_npz_array(str(tmp_path / "big.npz")), out, threads=4)
    assert np.array_equal(out, big)
    np.savez_compressed(tmp_path / "c.npz", data=big[:, :10], sf=1)
    with pytest.raises(ValueError, match="compressed"):
        SS.locate_npz_array(str(tmp_path / "c.npz"))


def test_tdt_tank_stream_reader_round_trip(tmp_path):
    """tdt_io: the .tsq / .tev stream stores the reference reads through tdt.read_block
    (ref: preprocess/io/tdt_blocks.py:6-18), on a synthetic tank in the TTank layout."""
    from decode_tonal_langauge_b200 import tdt_io
    rng = np.random.default_rng(1)
    ecog = rng.standard_normal((5, 4096)).astype(np.float32)
    audio = (rng.standard_normal((2, 8192)) * 1000).astype(np.int16)
    wide = rng.standard_normal((1, 512))
    blk_dir = tmp_path / "Sub1" / "HS1-B3"
    tdt_io.write_block(str(blk_dir), {"EOG1": (ecog, 3051.7578125), "ANIN": (audio, 24414.0625), "Wav8": (wide, 100.0)},
                       chunk=256)
    blk = tdt_io.read_block(str(blk_dir))
    assert np.array_equal(blk.streams.EOG1.data, ecog) and blk.streams.EOG1.data.dtype == np.float32
    assert blk.streams.EOG1.fs == pytest.approx(3051.7578125) and blk.streams.EOG1.channels == [1, 2, 3, 4, 5]
    assert np.array_equal(blk.streams.ANIN.data, audio) and blk.streams.ANIN.fs == pytest.approx(24414.0625)
    assert np.array_equal(blk.streams.Wav8.data, wide) and blk.info.stop > blk.info.start
    only = tdt_io.read_block(str(blk_dir), store="ANIN")
    assert list(vars(only.streams)) == ["ANIN"]
    # records out of time order and channels interleaved arbitrarily: sorted per channel by timestamp
    tsq = sorted(blk_dir.glob("*.tsq"))[0]
    heads = np.fromfile(tsq, dtype=tdt_io.TSQ_DTYPE)
    body = heads[2:-1].copy()
    rng.shuffle(body)
    np.concatenate([heads[:2], body, heads[-1:]]).tofile(tsq)
    assert np.array_equal(tdt_io.read_block(str(blk_dir)).streams.EOG1.data, ecog)
    # the stage driver's loader falls back to the native reader (no tdt wheel here)
    data = stages.load_block(str(blk_dir))
    assert np.array_equal(data["ecog"], ecog) and data["audio"].shape == (1, 8192) and data["audio_sf"] == pytest.approx(24414.0625)
    with pytest.raises(FileNotFoundError):
        tdt_io.read_block(str(tmp_path))
