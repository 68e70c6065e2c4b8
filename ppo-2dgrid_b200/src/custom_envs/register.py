"""Env-id registry (reference src/custom_envs/register.py:11-34).  Importing this module registers the five
`MERLIN-*-v0` ids; `make(env_id, **kwargs)` plays the role of `gym.make` (gymnasium is not a dependency)."""
from __future__ import annotations

import importlib

registry: dict = {}


def register(id, entry_point, **kwargs):
    registry[id] = (entry_point, kwargs)


def make(id, **kwargs):
    if id not in registry:
        raise KeyError(f"Environment `{id}` is not registered")
    entry_point, base = registry[id]
    mod, cls = entry_point.split(":")
    return getattr(importlib.import_module(mod), cls)(**{**base, **kwargs})


DIFFICULTY_OF = {}
for _id, _mod, _cls, _diff in (
    ("MERLIN-Easy-v0", "easy_env", "EasyEnv", "easy"),
    ("MERLIN-Medium-v0", "medium_env", "MediumEnv", "medium"),
    ("MERLIN-MediumHard-v0", "medium_hard_env", "MediumHardEnv", "mediumhard"),
    ("MERLIN-Hard-v0", "hard_env", "HardEnv", "hard"),
    ("MERLIN-Hardest-v0", "hardest_env", "HardestEnv", "hardest"),
):
    register(id=_id, entry_point=f"src.custom_envs.{_mod}:{_cls}")
    DIFFICULTY_OF[_id] = _diff
