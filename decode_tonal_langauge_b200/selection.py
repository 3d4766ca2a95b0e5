"""Discriminative / active channel selection on the device.

ref: channel_selection/discriminative.py:16-58,93-182, channel_selection/active.py:15-84,
channel_selection/utils.py:4-76.  ``run(data, params) -> {"selected_channels", "max_lengths",
"p_values"}`` for both modules; ``data`` is any mapping (an open npz) of numpy arrays or CUDA
tensors.  The per-(channel, timepoint) ANOVA and the longest-run scan run in
``ecog_anova_f`` / ``ecog_sig_runlength``; only the (C,) run lengths come back.
"""
from __future__ import annotations

from typing import Dict, Mapping

import numpy as np
import torch

from . import ops
from . import runtime as rt


def _epochs(data, name):
    try:
        rec = data[name]
    except KeyError:
        keys = list(data.keys()) if hasattr(data, "keys") else []
        raise KeyError(f"Recording '{name}' not found in data.Available keys: {keys}")
    return rec if rt.is_device(rec) else rt.to_device(np.asarray(rec))


def _sf(data, key):
    try:
        sf = data[key]
    except KeyError:
        raise ValueError("ECoG sampling frequency (ecog_sf) not found in the data.")
    return sf.item() if hasattr(sf, "item") else sf


def test_discriminative_power(data: Mapping, params: dict) -> Dict[str, np.ndarray]:
    """ref: discriminative.py:93-182; returns {'f_stat', 'p_value'} as (C, L) float64 numpy."""
    name = params.get("recording_name", "ecog")
    target = params["target"] if "target" in params else params["label"]      # Appendix B3
    series = _epochs(data, name)
    if series.dim() != 3:
        raise ValueError(f"Recording '{name}' must be a 3D array (n_samples, n_channels, n_timepoints).")
    try:
        labels = np.asarray(data[target]).squeeze()
    except KeyError:
        keys = list(data.keys()) if hasattr(data, "keys") else []
        raise KeyError(f"Labels '{target}' not found in data.Available keys: {keys}")
    if labels.ndim != 1:
        raise ValueError(f"Labels '{target}' must be a 1D array (n_samples,) or 2D array with shape "
                         "(1, n_samples) or (n_samples, 1).")
    if labels.shape[0] != series.shape[0]:
        raise ValueError(f"Number of samples in '{target}' ({labels.shape[0]}) does not match number of "
                         f"samples in '{name}' ({series.shape[0]}).")
    if not np.issubdtype(labels.dtype, np.integer):
        raise ValueError(f"Labels for '{target}' must be integers.")
    _, groups = np.unique(labels, return_inverse=True)       # group order = np.unique order (:161,176)
    F, P = ops.anova_f(series, groups)
    return {"f_stat": F, "p_value": P}


test_discriminative_power.__test__ = False      # not a pytest test


def _select(P: torch.Tensor, threshold: float, length_threshold: int, *also: torch.Tensor):
    """Longest significant run per channel -> selected list; `also` are further device results wanted on
    the host (p-values, F): everything comes back with one synchronisation."""
    host = ops.to_host_many(ops.sig_runlength(P, threshold), *also)
    runs = host[0]
    sel = [int(c) for c in np.nonzero(runs > length_threshold)[0]]       # strict '>' (utils.py:73)
    return sel, runs, host[1:]


def discriminative_run(data: Mapping, params: dict) -> dict:
    """ref: discriminative.py:16-58."""
    p_threshold = params.get("p_threshold", 0.05)
    name = params.get("recording_name", "ecog")
    target = params["target"] if "target" in params else params["label"]
    sf = _sf(data, f"{name}_sf")
    res = test_discriminative_power(data, params)
    P = res["p_value"]
    sel, _, (p_host, f_host) = _select(P, p_threshold / P.shape[1], int(params["active_time_threshold"] * sf),
                                       P, res["f_stat"])
    print(f'Found {len(sel)} discriminative channels for target "{target}"')
    # max_lengths is always empty in the reference (utils.py:65-75 never appends)
    return {"selected_channels": sel, "max_lengths": [], "p_values": p_host, "f_stat": f_host}


def active_run(data: Mapping, params: dict) -> dict:
    """ref: active.py:15-84: two-group ANOVA rest vs ERP; ``p_values`` is the LAST channel's row
    and ``max_lengths`` lists the selected channels' runs, as in the reference (:72-84)."""
    erp_name = params.get("erp_name", "ecog")
    rest_name = params.get("rest_name", "ecog_rest")
    sf = _sf(data, "ecog_sf")
    length_threshold = int(params["active_time_threshold"] * sf)
    rest = _epochs(data, rest_name)
    erp = _epochs(data, erp_name)
    if tuple(erp.shape[1:2]) != tuple(rest.shape[1:2]):
        raise ValueError(f"Shape mismatch between '{erp_name}' and '{rest_name}': "
                         f"{tuple(erp.shape[1:2])} vs {tuple(rest.shape[1:2])}.")
    groups = np.r_[np.zeros(rest.shape[0], np.int32), np.ones(erp.shape[0], np.int32)]
    _, P = ops.anova_f(rest, groups, erp)                    # f_oneway(rest, erp) group order (:62)
    sel, runs, (p_host,) = _select(P, params["p_threshold"] / rest.shape[2], length_threshold, P)
    print(f"Found {len(sel)} active channels.")
    return {"selected_channels": sel, "max_lengths": [int(runs[c]) for c in sel],
            "p_values": p_host[-1].copy(), "p_values_all": p_host}
