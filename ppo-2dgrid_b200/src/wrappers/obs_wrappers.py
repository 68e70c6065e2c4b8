"""Observation wrappers the reference takes from minigrid (src/scenario_creator/scenario_creator.py:45-50):
RGBImgPartialObsWrapper swaps the symbolic 7x7x3 image for the 56x56x3 egocentric RGB frame, ImgObsWrapper
drops the dict.  The frame itself was already produced by the step kernel; these only select it."""
from __future__ import annotations

import numpy as np

from ..spaces import Box
from .core import Wrapper


class RGBImgPartialObsWrapper(Wrapper):
    def __init__(self, env, tile_size=8):
        super().__init__(env)
        if tile_size != 8:
            raise ValueError("the CUDA renderer is built for tile_size 8 (the reference's default)")
        self.tile_size = tile_size
        n = env.unwrapped.agent_view_size * tile_size
        self.observation_space = dict(env.observation_space)
        self.observation_space["image"] = Box(0, 255, (n, n, 3), np.uint8)

    def observation(self, obs):
        return {**obs, "image": self.unwrapped.get_frame(tile_size=self.tile_size, agent_pov=True)}

    def reset(self, **kwargs):
        obs, info = self.env.reset(**kwargs)
        return self.observation(obs), info

    def step(self, action):
        obs, r, te, tr, info = self.env.step(action)
        return self.observation(obs), r, te, tr, info


class FullyObsWrapper(Wrapper):
    """minigrid FullyObsWrapper (`observation.fully_observable: true`): the whole grid as `Grid.encode()` with the
    agent's cell marked (10, 0, dir); produced by merlin_env_full_obs."""

    def __init__(self, env):
        super().__init__(env)
        u = env.unwrapped
        self.observation_space = dict(env.observation_space)
        self.observation_space["image"] = Box(0, 255, (u.width, u.height, 3), np.uint8)

    def observation(self, obs):
        return {**obs, "image": self.unwrapped.gen_full_obs()}

    def reset(self, **kwargs):
        obs, info = self.env.reset(**kwargs)
        return self.observation(obs), info

    def step(self, action):
        obs, r, te, tr, info = self.env.step(action)
        return self.observation(obs), r, te, tr, info


class ImgObsWrapper(Wrapper):
    def __init__(self, env):
        super().__init__(env)
        self.observation_space = env.observation_space["image"]

    def reset(self, **kwargs):
        obs, info = self.env.reset(**kwargs)
        return obs["image"], info

    def step(self, action):
        obs, r, te, tr, info = self.env.step(action)
        return obs["image"], r, te, tr, info


class FlattenObservation(Wrapper):
    def __init__(self, env):
        super().__init__(env)
        shape = env.observation_space.shape
        self.observation_space = Box(0, 255, (int(np.prod(shape)),), env.observation_space.dtype)

    def reset(self, **kwargs):
        obs, info = self.env.reset(**kwargs)
        return np.asarray(obs).reshape(-1), info

    def step(self, action):
        obs, r, te, tr, info = self.env.step(action)
        return np.asarray(obs).reshape(-1), r, te, tr, info
