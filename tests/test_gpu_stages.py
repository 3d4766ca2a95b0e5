"""Contract test on the B200: the reference's YAML (with its shipped spellings) drives the
whole preprocess -> sample_collection -> channel_selection pipeline on synthetic block files;
outputs are compared file by file with the oracle."""
import json
import os

import numpy as np
import pytest
import yaml

from conftest import max_rel

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _textgrid(onsets, tone, syl, dur):
    rows, t = [], 0.0
    for o, tn, sy in zip(onsets, tone, syl):
        start = o + 0.2
        if start > t:
            rows.append((t, start, ""))
        rows.append((start, start + 0.05, f"{tn}{'ia'[sy]}"))
        t = start + 0.05
    rows.append((t, dur, ""))
    out = ['File type = "ooTextFile"', 'Object class = "TextGrid"', "", "xmin = 0", f"xmax = {dur}",
           "tiers? <exists>", "size = 1", "item []:", "    item [1]:", '        class = "IntervalTier"',
           '        name = "success"', "        xmin = 0", f"        xmax = {dur}",
           f"        intervals: size = {len(rows)}"]
    for i, (a, b, m) in enumerate(rows, 1):
        out += [f"        intervals [{i}]:", f"            xmin = {a}", f"            xmax = {b}",
                f'            text = "{m}"']
    return "\n".join(out) + "\n"


def test_yaml_pipeline_end_to_end(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from decode_tonal_langauge_b200 import stages, synth, install_dropin
    from oracle import chains as OC, epochs as OE, selection as OS

    install_dropin()
    fs, C, dur = 2000, 16, 120.0
    T = int(fs * dur)
    raw = tmp_path / "raw" / "Sub1" / "tdt"
    x, (onsets, tone, syl) = synth.session(9, C, T, fs, n_events=80)
    aud = synth.audio(9, dur, 2441.40625)
    blk = raw / "HS1-B1"
    blk.mkdir(parents=True)
    np.savez(blk / "ecog.npz", data=x, sf=fs)
    np.savez(blk / "audio.npz", data=aud, sf=2441.40625)
    tg_dir = tmp_path / "tg" / "subject_1"
    tg_dir.mkdir(parents=True)
    (tg_dir / "HS1_B1.TextGrid").write_text(_textgrid(onsets, tone, syl, dur))

    steps = [
        {"module": "preprocess.downsample", "params": {"downsample_freq": 400}},
        {"module": "preprocess.frequency_filter", "params": {"bands": [
            {"method": "hilbert", "params": {"freq_ranges": [70, 150], "envelope": True}},
            {"method": "butter", "params": {"freqs": [0.3, 100], "filter_type": "bandpass"}}]}},
        {"module": "preprocess.zscore_rereference", "params": {"rereference_interval": [0.0, 25.0]}}]
    cfg = {
        "preprocess": {"module": "preprocess_main", "params": {
            "pipeline": {"module": "preprocess.pipelines.subject_block",
                         "params": {"subject_dirs": ["Sub1/tdt"], "subject_ids": [1]}},
            "io": {"module": "preprocess.io.tdt_blocks",
                   "params": {"root_dir": str(tmp_path / "raw"), "output_dir": str(tmp_path / "processed")}},
            "preprocessor": {"module": "preprocess.preprocessor"},
            "modalities": {"ecog": {"type": "signal", "preprocessing": {"steps": steps}},
                           "audio": {"type": "signal"}}}},
        "sample_collection": {"module": "extract_samples", "params": {
            "io": {"output_dir": str(tmp_path / "samples"), "textgrid_root": str(tmp_path / "tg")},
            "subjects": {1: {"start_offset": 0.2, "tier_list": ["success"], "blocks": [1],
                             "textgrid_dir": "subject_1", "rest_period": [0.0, 25.0], "sample_length": 1.0}},
            "settings": {"syllable_identifiers": ["i", "a"]}}},
        "channel_selection": {"module": "channel_selection_main", "params": {
            "io": {"output_dir": str(tmp_path / "sel")},
            "selections": [
                {"module": "channel_selection.active", "selection_name": "active_channels",
                 "params": {"p_threshold": 0.01, "active_time_threshold": 0.1, "rest_name": "ecog_rest",
                            "erp_name": "ecog"}},
                {"module": "channel_selection.discriminative", "selection_name": "tone_discriminative",
                 "params": {"p_threshold": 0.01, "active_time_threshold": 0.1, "label": "tone",
                            "recording_name": "ecog"}}]}},
    }
    path = tmp_path / "cfg.yaml"
    path.write_text(yaml.dump(cfg))
    outputs = stages.run_pipeline(str(path))

    # ---- preprocess stage: directory name, provenance, block files
    setup = outputs["preprocess"]
    on_disk = yaml.safe_load(path.read_text())      # the name hashes the parameter reprs in FILE order
    assert os.path.basename(setup) == stages.generate_setup_name(on_disk["preprocess"]["params"]["modalities"])
    assert "preprocess" in yaml.safe_load(open(os.path.join(setup, "config.yaml")))
    ecog = np.load(os.path.join(setup, "subject_1", "B1_ecog.npz"))
    audio = np.load(os.path.join(setup, "subject_1", "B1_audio.npz"))
    assert int(ecog["sf"]) == 400 and ecog["data"].shape == (2 * C, int(T / 5)) and ecog["data"].dtype == np.float64
    assert np.array_equal(audio["data"], aud)
    ref, _ = OC.run_chain(x, fs, [{**s, "params": {**s["params"], **({"bands": [
        {"method": "hilbert", "params": {"freq_ranges": [70.0, 150.0], "envelope": True}}, s["params"]["bands"][1]]}
        if "bands" in s["params"] else {})}} for s in steps])
    assert max_rel(ecog["data"], ref) < 1e-5

    # ---- sample collection: epochs are bit copies of the stored block, indices as the oracle computes them
    sdir = outputs["sample_collection"]
    merged = yaml.safe_load(open(os.path.join(sdir, "config.yaml")))
    assert set(merged) == {"preprocess", "sample_collection"}
    smp = np.load(os.path.join(sdir, "subject_1.npz"))
    from decode_tonal_langauge_b200 import textgrid_io
    iv = textgrid_io.handle_textgrids(str(tg_dir), start_offset=0.2, tier_list=["success"], blocks=[1])
    want = OE.extract_epochs({1: {"start": iv[1]["start"].to_numpy(), "tone": iv[1]["tone"].to_numpy(),
                                  "syllable": list(iv[1]["syllable"])}},
                             {1: {"ecog": (ecog["data"], ecog["sf"][()]), "audio": (aud, audio["sf"][()])}},
                             ["i", "a"], 1.0, (0.0, 25.0))
    for key in ("ecog", "audio", "ecog_rest", "syllable", "tone"):
        assert np.array_equal(smp[key], want[key]), key
        assert smp[key].dtype == want[key].dtype, key
    assert smp["ecog"].shape[0] > 40

    # ---- channel selection: same selected sets as the oracle on the same epochs
    cdir = outputs["channel_selection"]
    got = json.load(open(os.path.join(cdir, "subject_1.json")))
    data = {k: smp[k] for k in smp.files}
    assert got["active_channels"] == OS.active(data, {"p_threshold": 0.01, "active_time_threshold": 0.1})[
        "selected_channels"]
    assert got["tone_discriminative"] == OS.discriminative(
        data, {"p_threshold": 0.01, "active_time_threshold": 0.1, "target": "tone"})["selected_channels"]
    assert set(yaml.safe_load(open(os.path.join(cdir, "config.yaml")))) == {
        "preprocess", "sample_collection", "channel_selection"}
    # re-running sample collection skips the existing subject file (reference behaviour)
    before = os.path.getmtime(os.path.join(sdir, "subject_1.npz"))
    cfg2 = yaml.safe_load(path.read_text())
    cfg2["sample_collection"]["params"]["io"]["recording_dir"] = setup
    assert stages.extract_samples_run(cfg2) == sdir
    assert os.path.getmtime(os.path.join(sdir, "subject_1.npz")) == before
