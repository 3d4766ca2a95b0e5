"""Minimal Praat TextGrid reader (interval tiers, long and short text formats) and the
reference's TextGrid -> interval-table logic (ref: data_loading/text_align.py:12-186).

The ``textgrid`` wheel the reference imports is not part of this image; only what
``read_textgrid`` touches is modelled: ``tiers[*].name`` and
``tiers[*].intervals[*].{minTime,maxTime,mark}``.
"""
from __future__ import annotations

import os
import re
import warnings
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import pandas as pd

from .epochs import extract_block_id


@dataclass
class Interval:
    minTime: float
    maxTime: float
    mark: str


@dataclass
class Tier:
    name: str
    intervals: List[Interval] = field(default_factory=list)


@dataclass
class TextGrid:
    tiers: List[Tier] = field(default_factory=list)

    @classmethod
    def fromFile(cls, path: str) -> "TextGrid":
        raw = open(path, "rb").read()
        for enc in ("utf-8-sig", "utf-16", "latin-1"):
            try:
                text = raw.decode(enc)
                break
            except UnicodeError:
                continue
        return cls.fromString(text)

    @classmethod
    def fromString(cls, text: str) -> "TextGrid":
        # tokenise into numbers and quoted strings; works for the long ("xmin = 0") and the
        # short (bare values) formats alike
        # (quoted strings are matched first so brackets inside marks survive; "item [3]:" style
        # indices of the long format are dropped)
        pattern = re.compile(r'"((?:[^"]|"")*)"|(\[[^\]"]*\])|(-?\d+(?:\.\d+)?(?:[eE][-+]?\d+)?)')
        toks = []
        for m in pattern.finditer(text):
            if m.group(2) is not None:
                continue
            if m.group(3) is not None:
                toks.append((float(m.group(3)), False))
            else:
                toks.append((m.group(1).replace('""', '"'), True))
        # header: "ooTextFile" "TextGrid" xmin xmax [<exists>] size
        pos = 0
        while pos < len(toks) and toks[pos][1]:
            pos += 1
        pos += 2                                        # xmin xmax
        n_tiers = int(toks[pos][0]); pos += 1
        grid = cls()
        for _ in range(n_tiers):
            kind = toks[pos][0]; name = toks[pos + 1][0]; pos += 2
            pos += 2                                    # tier xmin xmax
            count = int(toks[pos][0]); pos += 1
            tier = Tier(str(name))
            for _ in range(count):
                if kind == "IntervalTier":
                    lo, hi, mark = toks[pos][0], toks[pos + 1][0], toks[pos + 2][0]
                    pos += 3
                    tier.intervals.append(Interval(float(lo), float(hi), str(mark)))
                else:                                   # TextTier points: (time, mark) -> skipped
                    pos += 2
            if kind == "IntervalTier":
                grid.tiers.append(tier)
        return grid


def read_textgrid(tg: TextGrid, start_offset: float, end_offset: float,
                  tier_list: Optional[List[str]] = None) -> pd.DataFrame:
    """Marks that start with a digit are trials: ``<tone digit><syllable char>...``
    (ref: text_align.py:114-151).  Tier names are compared lower-cased on both sides
    (the reference lower-cases only one side, Appendix B11)."""
    wanted = None if tier_list is None else {t.lower() for t in tier_list}
    trials: List[dict] = []
    for tier in tg.tiers:
        if wanted is not None and tier.name.lower() not in wanted:
            continue
        for iv in tier.intervals:
            if not iv.mark or not iv.mark[0].isdigit():
                continue
            start = iv.minTime - start_offset
            end = iv.maxTime + end_offset
            if trials and start < trials[-1]["end"]:
                warnings.warn(
                    f"Overlapping intervals detected in tier '{tier.name}' at time {iv.minTime:.2f}; "
                    "skipping this interval ...")
                continue
            trials.append({"start": np.around(start, decimals=1), "end": np.around(end, decimals=1),
                           "syllable": iv.mark[1] if len(iv.mark) > 1 else "", "tone": int(iv.mark[0])})
    return pd.DataFrame(trials, columns=["start", "end", "syllable", "tone"])


def handle_textgrids(data_dir: str, start_offset: float = 0.0, end_offset: float = 0.0,
                     tier_list: Optional[List[str]] = None,
                     blocks: Optional[List[int]] = None) -> Dict[int, pd.DataFrame]:
    """One interval table per block (ref: text_align.py:12-80); files are visited in sorted
    order so the result does not depend on the filesystem."""
    out: Dict[int, pd.DataFrame] = {}
    for name in sorted(os.listdir(data_dir)):
        if not name.endswith(".TextGrid"):
            continue
        block = extract_block_id(name)
        if (blocks is not None and block not in blocks) or block in out:
            continue
        out[block] = read_textgrid(TextGrid.fromFile(os.path.join(data_dir, name)),
                                   start_offset, end_offset, tier_list)
    return out
