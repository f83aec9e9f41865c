"""Minimal stand-in for the diffusers ConfigMixin protocol the reference relies on
(`load_config`, `from_config`, `.config.<key>`; reference call sites: src/utils_models.py:158-182,
src/pipeline_conditional_ddim/pipeline_conditionial_ddim.py:45, src/utils_Img2Img.py:776-778).

The JSON files of models_configs/{denoiser,noise_scheduler}/ are parsed unchanged: keys starting with "_"
("_class_name", "_diffusers_version") and keys a constructor does not declare are ignored, as diffusers does.
"""
from __future__ import annotations

import inspect
import json
import os
from typing import Any, Dict


class FrozenConfig(dict):
    """dict with attribute access, immutable after construction (diffusers FrozenDict behaviour)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        raise AttributeError("config is frozen")

    def __setitem__(self, k, v):
        raise TypeError("config is frozen")


class ConfigMixin:
    config_name = "config.json"

    def register_to_config(self, **kwargs):
        object.__setattr__(self, "_internal_config", FrozenConfig(kwargs))

    @property
    def config(self) -> FrozenConfig:
        return self._internal_config

    @classmethod
    def load_config(cls, path_or_dict, **kwargs) -> Dict[str, Any]:
        """Accepts a dict, a JSON file, or a directory holding `cls.config_name` (optionally under `subfolder`)."""
        if isinstance(path_or_dict, dict):
            return dict(path_or_dict)
        path = str(path_or_dict)
        sub = kwargs.get("subfolder")
        if os.path.isdir(path):
            if sub:
                path = os.path.join(path, sub)
            path = os.path.join(path, cls.config_name)
        with open(path, "r", encoding="utf-8") as f:
            return json.load(f)

    @classmethod
    def _init_keys(cls):
        sig = inspect.signature(cls.__init__)
        return [k for k, p in sig.parameters.items()
                if k != "self" and p.kind not in (p.VAR_KEYWORD, p.VAR_POSITIONAL)]

    @classmethod
    def from_config(cls, config, **kwargs):
        cfg = dict(config)
        keys = cls._init_keys()
        init = {k: v for k, v in cfg.items() if k in keys and not k.startswith("_")}
        init.update({k: v for k, v in kwargs.items() if k in keys})
        return cls(**init)

    def save_config(self, save_directory: str):
        os.makedirs(save_directory, exist_ok=True)
        d = {"_class_name": type(self).__name__, **{k: (list(v) if isinstance(v, tuple) else v) for k, v in self.config.items()}}
        with open(os.path.join(save_directory, self.config_name), "w", encoding="utf-8") as f:
            json.dump(d, f, indent=2, sort_keys=True)
