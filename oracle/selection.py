"""Oracle restatement of discriminative / active channel selection.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).
ref: channel_selection/discriminative.py:16-58,93-182, channel_selection/active.py:15-84,
channel_selection/utils.py:4-76; ANOVA follows scipy: stats/_stats_py.py ``f_oneway``
(equal_var=True) -- centre on the grand mean, SST, SSB, F, p = fdtrc(dfb, dfw, F).
"""
from __future__ import annotations

from typing import Dict, List, Mapping, Sequence

import numpy as np
from scipy import special


def anova_f(groups: Sequence[np.ndarray]):
    """One-way ANOVA along axis 0 of each (n_g, ...) group; returns (F, p) with the
    trailing shape.  scipy: stats/_stats_py.py:3965-3992,4045-4052."""
    groups = [np.asarray(g) for g in groups]
    alldata = np.concatenate(groups, axis=0)
    N = alldata.shape[0]
    G = len(groups)
    is_const = [np.all(np.diff(g, axis=0) == 0, axis=0) for g in groups]
    all_const = np.all(np.stack(is_const, axis=0), axis=0)
    all_same = np.all(np.diff(alldata, axis=0) == 0, axis=0)
    offset = alldata.mean(axis=0, keepdims=True)
    centred = alldata - offset
    norm_ss = centred.sum(axis=0) ** 2.0 / N
    sstot = np.einsum("i...,i...->...", centred, centred) - norm_ss
    ssb = 0
    for g in groups:
        ssb = ssb + (g - offset).sum(axis=0) ** 2.0 / g.shape[0]
    ssb = ssb - norm_ss
    ssw = sstot - ssb
    dfb, dfw = G - 1, N - G
    with np.errstate(divide="ignore", invalid="ignore"):
        F = (ssb / dfb) / (ssw / dfw)
    F = np.where(all_const, np.inf, F)
    F = np.where(all_same, np.nan, F)
    return F, special.fdtrc(dfb, dfw, F)


def longest_run(indices: np.ndarray) -> int:
    """ref: channel_selection/utils.py:4-30 (longest stretch of consecutive ints)."""
    best = cur = 1
    for a, b in zip(indices[:-1], indices[1:]):
        cur = cur + 1 if b == a + 1 else 1
        best = max(best, cur)
    return best


def significant_channels(p: np.ndarray, p_threshold: float, length_threshold: int) -> List[int]:
    """ref: channel_selection/utils.py:62-76 (Bonferroni over L, strict '>')."""
    thr = p_threshold / p.shape[1]
    out = []
    for ch in range(p.shape[0]):
        idx = np.where(p[ch] < thr)[0]
        if len(idx) > 0 and longest_run(idx) > length_threshold:
            out.append(ch)
    return out


def discriminative(data: Mapping[str, np.ndarray], params: Mapping) -> Dict:
    """ref: discriminative.py:16-58,93-182."""
    name = params.get("recording_name", "ecog")
    target = params["target"]
    sf = data[f"{name}_sf"]
    series = data[name]
    labels = np.asarray(data[target]).squeeze()
    if series.ndim != 3:
        raise ValueError("recording must be (n_samples, n_channels, n_timepoints)")
    if labels.ndim != 1 or labels.shape[0] != series.shape[0]:
        raise ValueError("labels must be 1-D and match the number of samples")
    if not np.issubdtype(labels.dtype, np.integer):
        raise ValueError(f"Labels for '{target}' must be integers.")
    uniq = np.unique(labels)
    C, L = series.shape[1:]
    F = np.zeros((C, L))
    P = np.zeros((C, L))
    for ch in range(C):
        x = series[:, ch, :]
        F[ch], P[ch] = anova_f([x[labels == g, :] for g in uniq])
    sel = significant_channels(P, params.get("p_threshold", 0.05),
                               int(params["active_time_threshold"] * sf))
    return {"selected_channels": sel, "max_lengths": [], "p_values": P, "f_stat": F}


def active(data: Mapping[str, np.ndarray], params: Mapping) -> Dict:
    """ref: active.py:15-84.  ``p_values`` is the LAST channel's row only and
    ``max_lengths`` lists the selected channels' runs (active.py:72-84)."""
    erp = data[params.get("erp_name", "ecog")]
    rest = data[params.get("rest_name", "ecog_rest")]
    sf = data["ecog_sf"]
    length_threshold = int(params["active_time_threshold"] * sf)
    if erp.shape[1:2] != rest.shape[1:2]:
        raise ValueError("Shape mismatch between ERP and rest recordings.")
    thr = params["p_threshold"] / rest.shape[2]
    sel, runs, p = [], [], None
    for ch in range(rest.shape[1]):
        _, p = anova_f([rest[:, ch, :], erp[:, ch, :]])
        idx = np.where(p < thr)[0]
        if len(idx) == 0:
            continue
        run = longest_run(idx)
        if run > length_threshold:
            sel.append(ch)
            runs.append(run)
    return {"selected_channels": sel, "max_lengths": runs, "p_values": p}
