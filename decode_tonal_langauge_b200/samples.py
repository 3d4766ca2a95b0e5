"""Consumer of the hot path's outputs: ``subject_<id>.npz`` + ``subject_<id>.json`` -> selected-channel
epochs and joint labels, resident on the device.

ref: data_loading/sample_loading.py:9-139 (``ClassificationSampleHandler``): same constructor
parameters (``sample_path``, ``channel_file``, ``targets``, ``features``), same ``load_data`` keys,
same joint label coding and channel union, same errors.  ``features`` comes back as a float CUDA
tensor (the channel filter runs in ``ecog_channel_select``); pass ``as_numpy=True`` for the
reference's numpy array.  The classifier models and their training are out of scope.
"""
from __future__ import annotations

import json
from argparse import Namespace
from typing import List, Optional

import numpy as np
import torch

from . import ops
from . import runtime as rt


class ClassificationSampleHandler:
    def __init__(self, params: Namespace):
        self.sample_path = params.sample_path
        self.channel_file = params.channel_file if hasattr(params, "channel_file") else None
        self.dataset = np.load(self.sample_path)
        self.channels = None
        self.targets = getattr(params, "targets", None)
        if isinstance(self.targets, str):
            self.targets = [self.targets]
        self.params = params

    def load_data(self, as_numpy: bool = False) -> dict:
        """ref: sample_loading.py:34-88."""
        try:
            features = self.dataset[self.params.features]
        except KeyError:
            raise KeyError(
                f"The dataset in {self.sample_path} does not contain {getattr(self.params, 'features', None)}. "
                f"Available keys: {', '.join(self.dataset.keys())}")
        target_labels, n_classes_dict = [], {}
        for target in self.targets:
            if target not in self.dataset:
                raise KeyError(f"The dataset does not contain '{target}' key. "
                               f"Available keys: {', '.join(self.dataset.keys())}")
            target_labels.append(self.dataset[target].flatten())
            n_classes_dict[target] = len(np.unique(self.dataset[target]))
        labels = np.zeros_like(target_labels[0], dtype=int)        # joint code, first target fastest (:68-72)
        multiplier = 1
        for target_label in target_labels:
            labels += target_label * multiplier
            multiplier *= len(np.unique(target_label))
        self.channels = self._filter_channels(features.shape[1])
        dev = rt.to_device(np.asarray(features), dtype=None)
        if dev.element_size() not in (4, 8):
            dev = dev.to(torch.float32)
        selected = ops.channel_select(dev, self.channels)
        return {"features": rt.to_host(selected) if as_numpy else selected, "labels": labels,
                "selected_channels": self.channels, "n_classes_dict": n_classes_dict}

    def _filter_channels(self, n_channels: int) -> np.ndarray:
        """ref: sample_loading.py:90-123: union of ``<target>_discriminative`` lists, sorted."""
        if self.channel_file is None:
            return np.arange(n_channels)
        with open(self.channel_file, "r") as f:
            channel_selections = json.load(f)
        channels = set()
        for target in self.targets:
            key = f"{target}_discriminative"
            if key not in channel_selections:
                raise KeyError(f"Channel selection for '{key}' not found in the file {self.channel_file}. "
                               f"Available keys: {', '.join(channel_selections.keys())}")
            channels.update(channel_selections[key])
        if not channels:
            raise ValueError(f"No channels found for the targets: {', '.join(self.targets)}. "
                             f"Please check the channel file {self.channel_file}")
        return np.array(sorted(channels))

    def prepare_torch_dataset(self, features, labels, device: str = "cuda"):
        """ref: sample_loading.py:126-139 (float32 features and labels on ``device``)."""
        from torch.utils.data import TensorDataset
        x = features if isinstance(features, torch.Tensor) else torch.as_tensor(np.asarray(features))
        return TensorDataset(x.to(device=device, dtype=torch.float32),
                             torch.as_tensor(np.asarray(labels), dtype=torch.float32).to(device))
