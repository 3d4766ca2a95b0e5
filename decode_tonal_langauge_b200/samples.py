"""Consumer of the hot path's outputs: ``subject_<id>.npz`` +
``subject_<id>.json`` -> selected-channel epochs resident on the device, plus joint class codes.

Contract taken from ref: data_loading/sample_loading.py:9-139 (``ClassificationSampleHandler``):

* constructor keys on the params Namespace: ``sample_path`` (npz), optional ``channel_file`` (the JSON
  written by the channel-selection stage), ``targets`` (one name or a list), ``features`` (npz key);
* ``load_data()`` returns ``features`` (events, selected channels, time), ``labels`` (one joint code per
  event: the first target varies fastest, radix = number of distinct values of each target),
  ``selected_channels`` and ``n_classes_dict``;
* the channel set is the sorted union of the ``<target>_discriminative`` lists of the JSON, or every
  channel without a channel file;
* missing npz keys / JSON entries raise ``KeyError``, an empty union raises ``ValueError``.

Written against that contract, not transliterated: the channel pick is a device gather
(``ecog_channel_select``) and ``features`` comes back as a CUDA tensor (``as_numpy=True`` for the
reference's numpy array).  The classifier models and their training are out of scope.
"""
from __future__ import annotations

import json
from argparse import Namespace
from typing import Dict, List, Sequence

import numpy as np
import torch

from . import ops
from . import runtime as rt


def joint_codes(columns: Sequence[np.ndarray]) -> np.ndarray:
    """Mixed-radix code of several label columns: ``sum_i column_i * prod_{j<i} n_distinct(column_j)``."""
    code = np.zeros(np.asarray(columns[0]).shape, dtype=int)
    radix = 1
    for col in columns:
        col = np.asarray(col)
        code = code + col * radix
        radix *= np.unique(col).size
    return code


def selected_union(selection: Dict[str, list], targets: Sequence[str], where: str = "the channel file") -> np.ndarray:
    """Sorted union of ``selection["<target>_discriminative"]`` over the targets."""
    picked = set()
    for name in targets:
        entry = f"{name}_discriminative"
        if entry not in selection:
            raise KeyError(f"Channel selection for '{entry}' not found in the file {where}. "
                           f"Available keys: {', '.join(selection)}")
        picked |= set(selection[entry])
    if len(picked) == 0:
        raise ValueError(f"No channels found for the targets: {', '.join(targets)}. "
                         f"Please check the channel file {where}")
    return np.asarray(sorted(picked))


class ClassificationSampleHandler:
    def __init__(self, params: Namespace):
        self.params = params
        self.sample_path = params.sample_path
        self.channel_file = getattr(params, "channel_file", None)
        targets = getattr(params, "targets", None)
        self.targets: List[str] = [targets] if isinstance(targets, str) else targets
        self.dataset = np.load(self.sample_path)
        self.channels = None

    def _column(self, key: str, what: str) -> np.ndarray:
        if key not in self.dataset:
            raise KeyError(f"The dataset in {self.sample_path} does not contain {what} '{key}'. "
                           f"Available keys: {', '.join(self.dataset.keys())}")
        return self.dataset[key]

    def load_data(self, as_numpy: bool = False) -> dict:
        features = self._column(getattr(self.params, "features", None), "the feature array")
        columns = [self._column(t, "the target").flatten() for t in self.targets]
        n_classes = {t: int(np.unique(self.dataset[t]).size) for t in self.targets}
        self.channels = self._filter_channels(features.shape[1])
        dev = rt.to_device(np.asarray(features), dtype=None)        # any dtype: the pick is a bit copy
        picked = ops.channel_select(dev, self.channels)
        return {"features": rt.to_host(picked) if as_numpy else picked, "labels": joint_codes(columns),
                "selected_channels": self.channels, "n_classes_dict": n_classes}

    def _filter_channels(self, n_channels: int) -> np.ndarray:
        if self.channel_file is None:
            return np.arange(n_channels)
        with open(self.channel_file, "r") as fh:
            selection = json.load(fh)
        return selected_union(selection, self.targets, self.channel_file)

    def prepare_torch_dataset(self, features, labels, device: str = "cuda"):
        """float32 features and float32 labels on ``device`` as a ``TensorDataset``."""
        from torch.utils.data import TensorDataset
        x = features if isinstance(features, torch.Tensor) else torch.from_numpy(np.asarray(features))
        y = torch.from_numpy(np.asarray(labels))
        return TensorDataset(x.to(device=device, dtype=torch.float32), y.to(device=device, dtype=torch.float32))
