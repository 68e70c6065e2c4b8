"""ThreeActionWrapper: the policy's {0, 1, 2} are MiniGrid's {left, right, forward}
(reference src/wrappers/three_action_wrapper.py:4-17).  For a batched CUDA env the 3-action table is the
kernel's default mode, so wrapping one only narrows `action_space`."""
from __future__ import annotations

import numpy as np

from ..spaces import Discrete
from .core import Wrapper


class ThreeActionWrapper(Wrapper):
    def __init__(self, env):
        super().__init__(env)
        self.action_space = Discrete(3)
        base = env.unwrapped.actions
        self._action_map = np.array([base.left, base.right, base.forward], dtype=np.int64)

    def action(self, act):
        return self._action_map[act]

    def step(self, action):
        return self.env.step(self.action(action))
