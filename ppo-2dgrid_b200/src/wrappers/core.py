"""Wrapper base: forwards everything to the wrapped env (what gymnasium.Wrapper gives the reference's wrappers)."""
from __future__ import annotations


class Wrapper:
    def __init__(self, env):
        self.env = env

    @property
    def unwrapped(self):
        return self.env.unwrapped

    def __getattr__(self, name):  # only reached when the wrapper itself lacks `name`
        if name.startswith("_") or name == "env":
            raise AttributeError(name)
        return getattr(self.env, name)

    def reset(self, **kwargs):
        return self.env.reset(**kwargs)

    def step(self, action):
        return self.env.step(action)

    def close(self):
        return self.env.close()
