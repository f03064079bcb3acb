"""Minimal ``ConfigMixin`` mirror: the parts of diffusers' config handling the reference reaches.

Reference call sites: ``model.config.sample_size`` / ``.in_channels`` inside ``DDPMPipeline.__call__``
(DriveSceneGen/scripts/generation.py:14), ``noise_scheduler.num_train_timesteps`` resolved through the config
``__getattr__`` fallback (DriveSceneGen/pipeline/training_pipeline.py:76), ``save_pretrained`` /
``from_pretrained`` (training_pipeline.py:107, generation.py:7, scripts/train.py:59).
"""
from __future__ import annotations

import json
import os
from collections import OrderedDict
from typing import Any, Dict

DIFFUSERS_VERSION = "0.20.0"


class FrozenDict(OrderedDict):
    """Read-only dict with attribute access (upstream ``configuration_utils.FrozenDict``)."""

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__frozen = True

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setitem__(self, key, value):
        if getattr(self, "_FrozenDict__frozen", False):
            raise Exception(f"You cannot use ``__setitem__`` on a {self.__class__.__name__} instance.")
        super().__setitem__(key, value)

    def __setattr__(self, name, value):
        if getattr(self, "_FrozenDict__frozen", False) and not name.startswith("_FrozenDict"):
            raise Exception(f"You cannot use ``__setattr__`` on a {self.__class__.__name__} instance.")
        super().__setattr__(name, value)


def _jsonable(v):
    if isinstance(v, tuple):
        return [_jsonable(x) for x in v]
    if isinstance(v, list):
        return [_jsonable(x) for x in v]
    return v


class ConfigMixin:
    config_name = "config.json"

    def register_to_config(self, **kwargs):
        object.__setattr__(self, "_internal_dict", FrozenDict(kwargs))

    @property
    def config(self) -> FrozenDict:
        return self._internal_dict

    def _config_json(self) -> Dict[str, Any]:
        d = {"_class_name": self.__class__.__name__, "_diffusers_version": DIFFUSERS_VERSION}
        d.update({k: _jsonable(v) for k, v in self.config.items()})
        return d

    def save_config(self, save_directory: str):
        os.makedirs(save_directory, exist_ok=True)
        with open(os.path.join(save_directory, self.config_name), "w", encoding="utf-8") as f:
            f.write(json.dumps(self._config_json(), indent=2, sort_keys=True) + "\n")

    @classmethod
    def load_config(cls, directory: str) -> Dict[str, Any]:
        with open(os.path.join(directory, cls.config_name), "r", encoding="utf-8") as f:
            d = json.load(f)
        return {k: v for k, v in d.items() if not k.startswith("_")}

    @classmethod
    def from_config(cls, config, **kwargs):
        import inspect
        cfg = dict(config)
        cfg.update(kwargs)
        accepted = set(inspect.signature(cls.__init__).parameters) - {"self"}
        # upstream ignores config entries the class does not take (e.g. DDPM -> DDIM scheduler swaps)
        return cls(**{k: v for k, v in cfg.items() if k in accepted})
