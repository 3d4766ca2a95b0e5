"""Event-locked epoch extraction (ref: data_loading/text_align.py:189-462).

Index arithmetic stays on the host in float64 -- ``int(start * sf)`` truncation decides
bit-exact indices -- and the gathers run on the device
(``ecog_epoch_gather``).  ``extract_epochs`` is the in-memory form (numpy or CUDA
sources); ``extract_ecog_audio`` keeps the reference's file-walking entry point.
"""
from __future__ import annotations

import os
import re
import warnings
from typing import Dict, List, Mapping, Optional, Sequence, Tuple

import numpy as np
import torch

from . import ops
from . import runtime as rt


def extract_block_id(filename: str) -> int:
    """ref: data_loading/utils.py:6-29."""
    m = re.search(r"B(\d+)", filename)
    if not m:
        raise ValueError(f"No block ID found in filename: {filename}")
    return int(m.group(1))


def match_filename(file: str, file_format: str, kwords: Optional[List[str]] = None) -> bool:
    """ref: data_loading/utils.py:82-116."""
    return file.endswith(file_format) and all(w in file for w in (kwords or []))


def onset_indices(starts: Sequence[float], sf, length: float) -> Tuple[np.ndarray, int]:
    """ref: text_align.py:291-292,381-382: ``int(row.start * sf)``, ``int(length * sf)``."""
    first = np.fromiter((int(s * sf) for s in starts), dtype=np.int64, count=len(starts))
    return first, int(length * sf)


def rest_indices(rest_period, earliest_start, sf, length, block=None) -> Tuple[np.ndarray, int]:
    """ref: text_align.py:313-340."""
    seg = int(length * sf)
    r0 = int(rest_period[0] * sf)
    r1 = int(rest_period[1] * sf)
    if rest_period[1] > earliest_start:
        warnings.warn(
            f"Rest period end ({rest_period[1]} s) is after the earliest interval start "
            f"for block {block} (earliest event time: {earliest_start} s). Reducing rest period end ...")
        r1 = int(earliest_start * sf)
    idx = [i for i in range(r0, r1, seg) if i + seg <= r1] if seg > 0 else []
    return np.asarray(idx, dtype=np.int64), seg


def syllable_codes(marks: Sequence[str], syllables: Sequence[str]) -> np.ndarray:
    """ref: text_align.py:308-311 (pd.Categorical codes: index in ``syllables`` or -1, int8)."""
    lut = {s: i for i, s in enumerate(syllables)}
    return np.fromiter((lut.get(str(m), -1) for m in marks), dtype=np.int8, count=len(marks))


def _column(table, name):
    col = table[name]
    return col.to_numpy() if hasattr(col, "to_numpy") else np.asarray(col)


def _gather(src, first: np.ndarray, n: int, what: str, block) -> torch.Tensor:
    """Device gather with the reference's overrun error (text_align.py:294-299)."""
    T = src.shape[1]
    bad = np.nonzero(first + n > T)[0]
    if bad.size:
        s = int(first[bad[0]])
        raise ValueError(
            f"Requested sample length exceeds {what} data length for block {block}. "
            f"Start: {s}, End: {s + n}; Data length: {T}.")
    d = src if rt.is_device(src) else rt.to_device(src, dtype=None)
    return ops.epoch_gather(d, first, n)


def extract_epochs(intervals: Mapping[int, Mapping], recordings: Mapping[int, Mapping[str, tuple]],
                   syllables: Sequence[str], length: float = 1.0,
                   rest_period: Optional[Sequence[float]] = None, to_numpy: bool = True) -> Dict:
    """In-memory epoch extraction.

    ``intervals[block]``: table with columns start / syllable / tone (DataFrame or dict);
    ``recordings[block]`` = {"ecog": (array (C,T), sf), "audio": (array (1,Ta), sf)} with numpy
    or CUDA arrays.  Blocks are merged in ascending block id (the reference follows
    ``os.listdir`` order, text_align.py:250,418-422)."""
    ecog_blocks = sorted(b for b in recordings if b in intervals and "ecog" in recordings[b])
    audio_blocks = sorted(b for b in recordings if b in intervals and "audio" in recordings[b])
    if ecog_blocks != audio_blocks:
        raise ValueError(
            "Mismatch between ECoG and audio samples blocks. "
            "Ensure both ECoG and audio files are present for each block."
            f" ECoG blocks found: {ecog_blocks}, Audio blocks found: {audio_blocks}.")
    if not ecog_blocks:
        raise ValueError(
            "No valid blocks found in the specified directories."
            f"Blocks in textgrids: {list(intervals.keys())}. ")
    erp, rest, aud, syl, tone = [], [], [], [], []
    ecog_sf = audio_sf = None
    for b in ecog_blocks:
        iv = intervals[b]
        starts = _column(iv, "start").astype(np.float64)
        ecog, ecog_sf = recordings[b]["ecog"]
        first, n = onset_indices(starts, ecog_sf, length)
        erp.append(_gather(ecog, first, n, "ECoG", b))
        tone.append(_column(iv, "tone"))
        syl.append(syllable_codes(list(_column(iv, "syllable")), syllables))
        if rest_period is not None:
            rfirst, seg = rest_indices(rest_period, starts.min(), ecog_sf, length, b)
            rest.append(_gather(ecog, rfirst, seg, "ECoG", b))
        audio, audio_sf = recordings[b]["audio"]
        afirst, an = onset_indices(starts, audio_sf, length)
        aud.append(_gather(audio[:1], afirst, an, "audio", b)[:, 0, :])
    tones = np.concatenate(tone, axis=0)
    if tones.size and tones.min() > 0:                     # ref :429-431
        tones = tones - tones.min()
    cat = lambda parts: parts[0] if len(parts) == 1 else torch.cat(parts, dim=0)
    out = {"ecog": cat(erp), "ecog_sf": ecog_sf, "audio": cat(aud), "audio_sf": audio_sf,
           "syllable": np.concatenate(syl, axis=0), "tone": tones}
    if rest_period is not None:
        out["ecog_rest"] = cat(rest)
    if to_numpy:
        for k in ("ecog", "audio", "ecog_rest"):
            if k in out:
                out[k] = rt.to_host(out[k])
    return out


def extract_ecog_audio(intervals, recording_dir: str, syllables: List[str], length: float = 1.0,
                       output_path: Optional[str] = None, rest_period: Optional[Tuple[float]] = None,
                       recording_format: str = "npz") -> Dict[str, np.ndarray]:
    """ref: data_loading/text_align.py:189-462.  Files ``B<block>_*ecog*.npz`` and
    ``B<block>_*sound*.npz`` (``*audio*`` also accepted: the reference's own writer uses that
    name, Appendix B5) with keys ``data`` and ``sf``."""
    print("Syllable mapping used: ", dict(enumerate(syllables)))
    recordings: Dict[int, dict] = {}
    for file in sorted(os.listdir(recording_dir)):
        if not file.endswith(recording_format):
            continue
        kind = "ecog" if "ecog" in file else ("audio" if ("sound" in file or "audio" in file) else None)
        if kind is None:
            continue
        block = extract_block_id(file)
        if block not in intervals:
            continue
        if kind in recordings.get(block, {}):
            warnings.warn(f"Found multiple {kind} files for block {block}, skipping file {file}. ")
            continue
        dataset = np.load(os.path.join(recording_dir, file))
        for key in ("data", "sf"):
            if key not in dataset:
                raise KeyError(
                    f"Expected key '{key}' not found in the npz file {file}. "
                    f"Existing keys {list(dataset.keys())}.")
        data, sf = dataset["data"], dataset["sf"]
        sf = sf[()] if getattr(sf, "ndim", 1) == 0 else sf
        print(f"{'ECoG' if kind == 'ecog' else 'Audio'} recording length for block {block}:",
              data.shape[1] / sf, " s")
        recordings.setdefault(block, {})[kind] = (data, sf)
    out = extract_epochs(intervals, recordings, syllables, length, rest_period, to_numpy=True)
    if rest_period is not None:
        print("ECoG rest samples shape:", out["ecog_rest"].shape)
    print("ECoG ERP samples shape:", out["ecog"].shape)
    print("Audio samples shape:", out["audio"].shape)
    if output_path is not None:
        np.savez(output_path, **out)
        print(f"ECoG and audio samples saved to {output_path}")
    return out
