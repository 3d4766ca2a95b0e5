"""YAML / Namespace / provenance helpers with the reference's observable behaviour
(ref: utils/config.py:8-84): same hash names, same merged ``config.yaml`` files."""
from __future__ import annotations

import hashlib
import json
import os
from argparse import Namespace
from typing import Any, Iterable, Optional

import yaml


def load_config(path: str) -> dict:
    with open(path, "r") as fh:
        return yaml.safe_load(fh)


def dict_to_namespace(obj: Any, exclude_keys: Optional[Iterable[str]] = None):
    """Nested dicts -> nested Namespaces; values under ``exclude_keys`` are kept verbatim
    (ref: utils/config.py:14-27)."""
    skip = set(exclude_keys or ())
    if isinstance(obj, dict):
        return Namespace(**{k: (v if k in skip else dict_to_namespace(v)) for k, v in obj.items()})
    if isinstance(obj, list):
        return [dict_to_namespace(v) for v in obj]
    return obj


def update_configuration(output_path: str, previous_config_path: str, new_module: str, new_module_cfg: dict) -> None:
    """Carry the upstream stage's config.yaml forward and add this stage's section (ref :58-71)."""
    merged = load_config(previous_config_path) if os.path.exists(previous_config_path) else {}
    if not os.path.exists(previous_config_path):
        print(f"Warning: config.yaml not found in {previous_config_path}")
    merged = merged or {}
    merged[new_module] = new_module_cfg
    with open(output_path, "w") as fh:
        yaml.dump(merged, fh)


def generate_hash_name_from_config(base_name: str, config: dict) -> str:
    """``<base>__<md5(json, sorted keys)[:6]>`` (ref :74-84)."""
    digest = hashlib.md5(json.dumps(config, sort_keys=True).encode()).hexdigest()
    return f"{base_name}__{digest[:6]}"


def yaml_hash_name(base_name: str, cfg: dict) -> str:
    """``<base>__<md5(yaml.dump(sorted))[:6]>`` used by sample collection (ref: extract_samples.py:136-144)."""
    digest = hashlib.md5(yaml.dump(cfg, sort_keys=True).encode()).hexdigest()
    return f"{base_name}__{digest[:6]}"
