"""Oracle restatement of event-locked epoch extraction.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).
Follows ref: data_loading/text_align.py:189-462 (``extract_ecog_audio``) with the
file walking removed: recordings arrive as in-memory arrays keyed by block.
"""
from __future__ import annotations

import warnings
from typing import Dict, Mapping, Optional, Sequence

import numpy as np


def onset_indices(starts, sf, length):
    """ref: text_align.py:291-292 / :381-382.  ``int(start * sf)`` truncates a
    float64 product toward zero; ``sf`` is whatever scalar the npz stored (int64
    400 after ``downsample``, float64 24414.0625 for raw audio)."""
    first = np.array([int(s * sf) for s in starts], dtype=np.int64)
    n = int(length * sf)
    return first, n


def rest_indices(rest_period, earliest_start, sf, length):
    """ref: text_align.py:313-340.  Full, non-overlapping segments of
    ``int(length*sf)`` samples from ``int(r0*sf)`` up to ``int(r1*sf)``, the end
    clipped to the earliest event onset (with a warning)."""
    seg = int(length * sf)
    r0 = int(rest_period[0] * sf)
    r1 = int(rest_period[1] * sf)
    if rest_period[1] > earliest_start:
        warnings.warn("Rest period end is after the earliest interval start; reducing.")
        r1 = int(earliest_start * sf)
    out = []
    for i in range(r0, r1, seg):
        if i + seg > r1:
            break
        out.append(i)
    return np.array(out, dtype=np.int64), seg


def gather(source: np.ndarray, first: np.ndarray, n: int, what="ECoG", block=None):
    """ref: text_align.py:294-304.  (N, C, n) stack of ``source[:, s:s+n]``."""
    T = source.shape[1]
    for s in first:
        if s + n > T:
            raise ValueError(
                f"Requested sample length exceeds {what} data length for block {block}. "
                f"Start: {s}, End: {s + n}; Data length: {T}.")
    if len(first) == 0:
        return np.zeros((0,) + (source.shape[0], n), dtype=source.dtype)
    return np.stack([source[:, s:s + n] for s in first], axis=0)


def syllable_codes(marks: Sequence[str], syllables: Sequence[str]) -> np.ndarray:
    """ref: text_align.py:308-311 -- ``pd.Categorical(..., categories=syllables).codes``:
    position in ``syllables`` or -1, int8."""
    lut = {s: i for i, s in enumerate(syllables)}
    return np.array([lut.get(m, -1) for m in marks], dtype=np.int8)


def extract_epochs(intervals: Mapping[int, Mapping[str, Sequence]],
                   recordings: Mapping[int, Mapping[str, tuple]],
                   syllables: Sequence[str], length: float = 1.0,
                   rest_period: Optional[Sequence[float]] = None) -> Dict[str, np.ndarray]:
    """ref: text_align.py:239-462.

    ``intervals[block]`` has columns ``start`` (float64 seconds, already rounded
    to one decimal, :145), ``syllable`` (str) and ``tone`` (int).
    ``recordings[block]`` = {"ecog": (data(C,T), sf), "audio": (data(1,Ta), sf)}.
    Blocks are merged in ascending block id (the reference uses ``os.listdir``
    order, :250,396,418-422, which is filesystem dependent).
    """
    blocks = sorted(b for b in recordings if b in intervals)
    if not blocks:
        raise ValueError("No valid blocks found in the specified directories.")
    erp, rest, aud, syl, tone = [], [], [], [], []
    ecog_sf = audio_sf = None
    for b in blocks:
        iv = intervals[b]
        starts = np.asarray(iv["start"], dtype=np.float64)
        if "ecog" not in recordings[b] or "audio" not in recordings[b]:
            raise ValueError("Mismatch between ECoG and audio samples blocks.")
        ecog, ecog_sf = recordings[b]["ecog"]
        first, n = onset_indices(starts, ecog_sf, length)
        erp.append(gather(ecog, first, n, "ECoG", b))
        tone.append(np.asarray(iv["tone"]))
        syl.append(syllable_codes(list(iv["syllable"]), syllables))
        if rest_period is not None:
            rfirst, seg = rest_indices(rest_period, starts.min(), ecog_sf, length)
            rest.append(gather(ecog, rfirst, seg, "ECoG", b))
        audio, audio_sf = recordings[b]["audio"]
        afirst, an = onset_indices(starts, audio_sf, length)
        aud.append(gather(audio[:1], afirst, an, "audio", b)[:, 0, :])
    tones = np.concatenate(tone, axis=0)
    if tones.min() > 0:                      # ref: text_align.py:429-431
        tones = tones - tones.min()
    out = {
        "ecog": np.concatenate(erp, axis=0),
        "ecog_sf": ecog_sf,
        "audio": np.concatenate(aud, axis=0),
        "audio_sf": audio_sf,
        "syllable": np.concatenate(syl, axis=0),
        "tone": tones,
    }
    if rest_period is not None:
        out["ecog_rest"] = np.concatenate(rest, axis=0)
    return out
