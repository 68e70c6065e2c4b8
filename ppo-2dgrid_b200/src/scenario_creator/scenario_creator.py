"""ScenarioCreator -- YAML scenario table -> environments (reference
src/scenario_creator/scenario_creator.py:10-73, same constructor, methods, attributes and errors).

`create_env` builds the reference's single-env wrapper stack (RGB partial obs -> image only -> optional
flatten -> three actions) over a CUDA-backed env; `create_batched_env` is the additive batched surface: N envs
of one difficulty stepped by one kernel launch, observations / rewards / flags as device tensors.
"""
from __future__ import annotations

import os

import numpy as np
import yaml

import src.custom_envs.register as _register  # noqa: F401  (registration side effect, as in the reference)
from src.wrappers.obs_wrappers import FlattenObservation, FullyObsWrapper, ImgObsWrapper, RGBImgPartialObsWrapper
from src.wrappers.three_action_wrapper import ThreeActionWrapper

DEFAULT_CONFIG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "config", "scenario.yaml")


class ScenarioCreator:
    def __init__(self, config_path: str = "src/config/scenario.yaml"):
        if not os.path.exists(config_path):
            if config_path == "src/config/scenario.yaml" and os.path.exists(DEFAULT_CONFIG):
                config_path = DEFAULT_CONFIG  # the packaged table, when not run from the repo root
            else:
                raise FileNotFoundError(f"Config not found: {config_path}")
        with open(config_path, "r") as f:
            self.config = yaml.safe_load(f)
        self.seed = self.config.get("seed", 42)
        self.global_cfg = self.config.get("global", {})
        self.obs_cfg = self.config.get("observation", {})
        self.rewards_cfg = self.config.get("rewards", {})
        self.logging_cfg = self.config.get("logging", {})
        self._validate_grid_sizes()

    def _validate_grid_sizes(self):
        sizes = set()
        for cfg in self.config["difficulties"].values():
            env_id = cfg["env_id"]
            if "-" in env_id and "x" in env_id:
                sizes.add(env_id.split("-")[-2])
        if len(sizes) > 1:
            raise ValueError(f"Multiple grid sizes detected: {sizes}")

    def _difficulty_cfg(self, difficulty):
        cfg = self.config["difficulties"].get(difficulty)
        if not cfg:
            raise ValueError(f"Unknown difficulty: {difficulty}")
        return cfg

    # ---- the reference's single-env path ----------------------------------------------------------------
    def create_env(self, difficulty: str = "easy", seed=None):
        """`seed` is accepted and ignored, as in the reference (the env is seeded by `reset(seed=...)`)."""
        cfg = self._difficulty_cfg(difficulty)
        env = _register.make(cfg["env_id"], **{**self.global_cfg, **cfg.get("params", {})})
        if self.obs_cfg.get("fully_observable", False):
            env = FullyObsWrapper(env)
        else:
            env = RGBImgPartialObsWrapper(env)
        env = ImgObsWrapper(env)
        if self.obs_cfg.get("flatten", False):
            env = FlattenObservation(env)
        return ThreeActionWrapper(env)

    def sample_scenarios(self, n: int = 5, difficulty: str = "easy"):
        return [self.create_env(difficulty) for _ in range(n)]

    # ---- batched surface (additive) ---------------------------------------------------------------------
    def create_batched_env(self, difficulty="easy", num_envs=1, device="cuda", layouts=None, seeds=None,
                           stuck_penalty=False, exploration_bonus=0.0, fomaml_mode=False, size=None,
                           want_symbolic=False, n_layouts=None, **env_kwargs):
        """N envs of `difficulty` on `device`.  Layout pool: `layouts=(cells u8[L, H*W], agent i32[L, 3])`, or
        generated on the host from `seeds` (one `reset(seed=s)` layout per seed; default: seeds 0..max(N, 1024)-1).
        `layouts="device"`: the pool (`n_layouts` entries, default max(N, 1024)) is generated on the GPU from the integer
        `seeds` -- fresh layouts of the right distribution, not the reference's per-seed layouts.
        `fomaml_mode`: finished envs restart on their own layout (src/fomaml.py:92) instead of the next one."""
        from merlin_b200 import BatchedMerlinEnv
        from merlin_b200 import layouts as _layouts

        cfg = self._difficulty_cfg(difficulty)
        params = {**self.global_cfg, **cfg.get("params", {})}
        size = int(size if size is not None else params.get("size", 16))
        diff = _register.DIFFICULTY_OF[cfg["env_id"]]
        common = dict(width=size, height=size)
        if isinstance(layouts, str):
            if layouts != "device":
                raise ValueError("layouts must be (cells, agent), None (host-generated from seeds) or 'device'")
            pool = dict(generate=(diff, int(seeds or 0), int(n_layouts or max(int(num_envs), 1024))))
            cells = agent = None
        else:
            if layouts is None:
                if seeds is None:
                    seeds = range(int(n_layouts or max(int(num_envs), 1024)))
                layouts = _layouts.generate(diff, size, seeds)
            cells, agent = (np.asarray(a) for a in layouts)
            pool = {}
        env = BatchedMerlinEnv(int(num_envs), cells, agent, **common, **pool,
                               max_steps=env_kwargs.pop("max_steps", params.get("max_steps")), device=device,
                               reset_mode="same" if fomaml_mode else "next", stuck_penalty=stuck_penalty,
                               exploration_bonus=exploration_bonus, want_symbolic=want_symbolic, **env_kwargs)
        env.difficulty, env.size = diff, size
        return env

    # ---- accessors --------------------------------------------------------------------------------------
    def get_env_id(self, difficulty: str) -> str:
        return self.config["difficulties"][difficulty]["env_id"]

    def get_logging_params(self) -> dict:
        return self.logging_cfg

    def get_observation_params(self) -> dict:
        return self.obs_cfg

    def get_env_size_str(self, difficulty: str) -> str:
        size = self.config["difficulties"][difficulty].get("params", {}).get("size", 16)
        return f"{size}x{size}"
