"""Stage drivers with the reference's entry points and on-disk contract:

    preprocess_main.main(config_path) / run(config)      ref: preprocess_main.py:8-27
    subject_block.run(...)                               ref: preprocess/pipelines/subject_block.py:54-103
    extract_samples.run(config) -> output dir            ref: extract_samples.py:16-123
    channel_selection_main.run(config) -> output dir     ref: channel_selection_main.py:19-92
    main.run_pipeline(config_path)                       ref: main.py:18-72

Directory names (md5 of the configuration), provenance ``config.yaml`` files, skip rules
and file names are the reference's; figures are not produced.
"""
from __future__ import annotations

import hashlib
import importlib
import json
import os
import warnings
from typing import Any, Dict, Iterator, Optional, Tuple

import numpy as np
import yaml

from . import config as cfgmod
from . import epochs, preprocessor, selection, textgrid_io

STAGES = ["preprocess", "sample_collection", "channel_selection", "training", "evaluation", "visualisation"]

# dotted names the YAML may use for the plug-ins that ship with this package
_BUILTIN = {
    "preprocess.preprocessor": "decode_tonal_langauge_b200.preprocessor",
    "preprocess.pipelines.subject_block": "decode_tonal_langauge_b200.stages",
    "preprocess.io.npz_blocks": "decode_tonal_langauge_b200.stages",
    "preprocess.io.tdt_blocks": "decode_tonal_langauge_b200.stages",
    "preprocess_main": "decode_tonal_langauge_b200.stages",
    "extract_samples": "decode_tonal_langauge_b200.stages",
    "channel_selection_main": "decode_tonal_langauge_b200.stages",
}


def _import(name: str):
    try:
        return importlib.import_module(name)
    except ModuleNotFoundError:
        if name in _BUILTIN:
            return importlib.import_module(_BUILTIN[name])
        raise


# ------------------------------------------------------------------ block IO
def load_block(block_path: str) -> dict:
    """Read one recording block.  A directory holding ``ecog.npz`` / ``audio.npz`` (keys ``data``,
    ``sf``) is read directly; a TDT tank block goes through ``tdt.read_block`` like the reference
    (ref: preprocess/io/tdt_blocks.py:6-18, streams EOG1 / ANIN) or, without that wheel, through the
    native stream reader ``tdt_io.read_block``."""
    if os.path.exists(os.path.join(block_path, "ecog.npz")):
        out = {}
        for key in ("ecog", "audio"):
            f = os.path.join(block_path, f"{key}.npz")
            if os.path.exists(f):
                z = np.load(f)
                out[key] = z["data"]
                out[f"{key}_sf"] = z["sf"][()]
        return out
    try:
        import tdt                                     # the vendor's reader when it is installed
        reader = getattr(tdt, "read_block", None)
    except ImportError:
        reader = None
    if reader is None:
        from . import tdt_io                           # stream stores read natively (.tsq / .tev)
        reader = tdt_io.read_block
    blk = reader(block_path)
    return {"ecog": blk.streams.EOG1.data, "audio": blk.streams.ANIN.data[:1, :],
            "ecog_sf": blk.streams.EOG1.fs, "audio_sf": blk.streams.ANIN.fs}


def save_block(setup_dir: str, subject_id: int, block_id: int, data_dict: dict) -> None:
    """``<setup>/subject_<id>/B<block>_<modality>.npz`` with keys data / sf (ref :21-35)."""
    out_dir = os.path.join(setup_dir, f"subject_{subject_id}")
    os.makedirs(out_dir, exist_ok=True)
    for key, value in data_dict.items():
        if key.endswith("_sf"):
            continue
        path = os.path.join(out_dir, f"B{block_id}_{key}.npz")
        np.savez(path, data=value, sf=data_dict.get(f"{key}_sf"))
        print(f"Saved {key} data to: {path}")


# ------------------------------------------------------- preprocess pipeline
def get_block_id(dirname: str) -> Optional[int]:
    tail = dirname.split("-")[-1].replace("B", "")
    try:
        return int(tail)
    except ValueError:
        print(f"Skipping directory '{dirname}' as it does not match expected format.",
              "Expected format: 'HS<subject_id>-<block_id>'.")
        return None


def iter_blocks(root_dir: str, subject_dirs, subject_ids=None) -> Iterator[Tuple[int, int, str]]:
    ids = subject_ids if subject_ids is not None else list(range(1, len(subject_dirs) + 1))
    for sid, sdir in zip(ids, subject_dirs):
        base = os.path.join(root_dir, sdir)
        for name in sorted(os.listdir(base)):            # sorted: deterministic across filesystems
            bid = get_block_id(name)
            if bid is not None:
                yield sid, bid, os.path.join(base, name)


def generate_setup_name(modalities_cfg: Dict[str, Any]) -> str:
    """``<step names joined by __>_<md5[:6]>`` over module names and param reprs (ref :42-51)."""
    steps = [s for m in modalities_cfg.values() for s in (m.get("preprocessing") or {}).get("steps", [])]
    if not steps:
        return "raw"
    readable = "__".join(s["module"].split(".")[-1] for s in steps)
    blob = "_".join(f"{s['module']}_{s.get('params', {})}" for s in steps)
    return f"{readable}_{hashlib.md5(blob.encode()).hexdigest()[:6]}"


def run(*args, **kwargs):
    """Dispatch on the call shape: stage entry ``run(config)`` for preprocess_main, or the
    pipeline plug-in ``run(pipeline_params, io_params, io_module, preprocessor_module, modalities_cfg)``."""
    if len(args) == 1 and isinstance(args[0], dict) and not kwargs:
        return preprocess_run(args[0])
    return subject_block_run(*args, **kwargs)


def subject_block_run(pipeline_params, io_params, io_module, preprocessor_module, modalities_cfg) -> str:
    setup_dir = os.path.join(io_params.output_dir, generate_setup_name(modalities_cfg))
    os.makedirs(os.path.join(setup_dir, "figures"), exist_ok=True)
    with open(os.path.join(setup_dir, "config.yaml"), "w") as fh:
        yaml.dump({"preprocess": {"pipeline": vars(pipeline_params), "io": vars(io_params),
                                  "modalities": modalities_cfg}}, fh)
    for sid, bid, path in iter_blocks(io_params.root_dir, pipeline_params.subject_dirs,
                                      getattr(pipeline_params, "subject_ids", None)):
        print(f"Processing block {bid} of subject {sid}...")
        data = io_module.load_block(path)
        params = cfgmod.dict_to_namespace({**vars(io_params), "block_id": bid, "subject_id": sid},
                                          exclude_keys=["root_dir", "output_dir"])
        preprocessor_module.preprocess_modalities(data, modalities_cfg, params, figure_dir=None)
        io_module.save_block(setup_dir, sid, bid, data)
    return setup_dir


def preprocess_run(config: dict) -> str:
    """``run(config)`` the stage runner needs (ref: main.py:28-47; the reference only ships
    ``main(config_path)``, Appendix B4)."""
    pre = (config.get("preprocess") or {}).get("params", {})
    pipe, io = pre.get("pipeline", {}), pre.get("io", {})
    prep = pre.get("preprocessor", {"module": "preprocess.preprocessor"})
    pipeline_module = _import(pipe.get("module", "preprocess.pipelines.subject_block"))
    io_module = _import(io.get("module", "preprocess.io.npz_blocks"))
    prep_module = _import(prep.get("module", "preprocess.preprocessor"))
    runner = getattr(pipeline_module, "subject_block_run", None) or pipeline_module.run
    return runner(cfgmod.dict_to_namespace(pipe.get("params", {})), cfgmod.dict_to_namespace(io.get("params", {})),
                  io_module, prep_module, pre.get("modalities", {}))


def main(config_path: str) -> None:
    preprocess_run(cfgmod.load_config(config_path))


# ------------------------------------------------------------ sample collection
def extract_samples_run(config: dict) -> str:
    section = config.get("sample_collection", {})
    pc = section.get("params", {})
    flat = {}
    for part in ("io", "settings"):
        flat.update(pc.get(part, {}))
    params = cfgmod.dict_to_namespace(flat)
    overwrite = getattr(params, "overwrite", False)
    out_dir = os.path.join(params.output_dir, cfgmod.yaml_hash_name(os.path.basename(params.recording_dir), section))
    os.makedirs(os.path.join(out_dir, "figures"), exist_ok=True)
    cfgmod.update_configuration(os.path.join(out_dir, "config.yaml"),
                                os.path.join(params.recording_dir, "config.yaml"), "sample_collection", section)
    for sid, sp in (pc.get("subjects") or {}).items():
        rec_dir = os.path.join(params.recording_dir, f"subject_{sid}")
        target = os.path.join(out_dir, f"subject_{sid}.npz")
        tg_dir = os.path.join(params.textgrid_root, sp["textgrid_dir"])
        if not os.path.exists(rec_dir):
            print(f"Recording directory {rec_dir} not found. Skipping...")
            continue
        if os.path.exists(target) and not overwrite:
            print(f"Output file {target} already exists. Skipping ...")
            continue
        if not os.path.exists(tg_dir):
            print(f"TextGrid directory {tg_dir} not found. Skipping...")
            continue
        intervals = textgrid_io.handle_textgrids(tg_dir, start_offset=sp.get("start_offset", 0.0),
                                                 tier_list=sp.get("tier_list"), blocks=sp.get("blocks"))
        if not intervals:
            raise ValueError("No intervals found in the TextGrid files. Check the directory and file naming "
                             f"conventions. Target blocks: {sp.get('blocks') or 'all'}")
        rest = sp.get("rest_period")
        epochs.extract_ecog_audio(intervals, rec_dir, syllables=params.syllable_identifiers,
                                  length=sp["sample_length"], output_path=target,
                                  rest_period=tuple(rest) if rest is not None else None)
    return out_dir


# ------------------------------------------------------------ channel selection
_SELECTORS = {"channel_selection.discriminative": selection.discriminative_run,
              "channel_selection.active": selection.active_run}


def channel_selection_run(config: dict) -> str:
    section = config.get("channel_selection", {})
    pc = section.get("params", {})
    io = cfgmod.dict_to_namespace(pc.get("io", {}))
    out_dir = os.path.join(io.output_dir,
                           cfgmod.generate_hash_name_from_config(os.path.basename(io.sample_dir), section))
    os.makedirs(os.path.join(out_dir, "figures"), exist_ok=True)
    cfgmod.update_configuration(os.path.join(out_dir, "config.yaml"), os.path.join(io.sample_dir, "config.yaml"),
                                "channel_selection", section)
    for name in sorted(os.listdir(io.sample_dir)):
        if not (name.startswith("subject_") and name.endswith(".npz")):
            continue
        sid = name.split("_")[1].split(".")[0]
        data = np.load(os.path.join(io.sample_dir, name))
        result = {}
        for sel in pc.get("selections", []):
            mod_name = sel["module"]
            fn = _SELECTORS.get(mod_name.replace("decode_tonal_langauge_b200.dropin.", "")) \
                or importlib.import_module(mod_name).run
            res = fn(data, sel.get("params", {}))
            result[sel["selection_name"]] = [int(c) for c in res["selected_channels"]]
            if not result[sel["selection_name"]]:
                warnings.warn(f"No active channels found for selection {sel['selection_name']} in subject {sid}.")
        with open(os.path.join(out_dir, f"subject_{sid}.json"), "w") as fh:
            json.dump(result, fh, indent=4)
        print(f"Saved results for subject {sid} to {os.path.join(out_dir, f'subject_{sid}.json')}.")
    return out_dir


# ------------------------------------------------------------------ stage loop
def _wire_io(outputs: Dict[str, str], stage: str, stage_cfg: dict) -> None:
    """Feed a stage's output directory into the next stage's io block (ref: main.py:55-72)."""
    io = stage_cfg.setdefault("params", {}).setdefault("io", {})
    if stage == "sample_collection" and "preprocess" in outputs:
        io.setdefault("recording_dir", outputs["preprocess"])
    elif stage in ("channel_selection", "training") and "sample_collection" in outputs:
        io.setdefault("sample_dir", outputs["sample_collection"])
    if stage == "training" and "channel_selection" in outputs:
        io.setdefault("channel_selection_dir", outputs["channel_selection"])


_STAGE_FUNCS = {"preprocess_main": preprocess_run, "extract_samples": extract_samples_run,
                "channel_selection_main": channel_selection_run}


def run_pipeline(config_path: str) -> Dict[str, str]:
    config = cfgmod.load_config(config_path)
    outputs: Dict[str, str] = {}
    for stage in STAGES:
        sc = config.get(stage)
        if not sc or sc.get("module") is None:
            continue
        print("----------- Running stage:", stage, "-----------")
        _wire_io(outputs, stage, sc)
        config[stage] = sc
        fn_name = sc.get("function", "run")
        if sc["module"] in _STAGE_FUNCS and fn_name == "run":
            fn = _STAGE_FUNCS[sc["module"]]
        else:
            module = importlib.import_module(sc["module"])
            if not hasattr(module, fn_name):
                raise ImportError(f"Module '{sc['module']}' does not have a function '{fn_name}'")
            fn = getattr(module, fn_name)
        result = fn(config)
        if isinstance(result, str):
            outputs[stage] = result
    return outputs
