"""StuckPenaltyWrapper (reference src/wrappers/stuck_penalty_wrapper.py:3-58): once the agent's position has
been unchanged for `max_stay` consecutive steps, every further such step adds `penalty`; `info["stuck"]`
reports it.  The counter, the comparison and the float64 reward add run inside the step kernel
(MERLIN_F_STUCK_PENALTY) -- this class switches that mode on for the wrapped env and keeps the reference's
attributes (`max_stay`, `penalty`, `stay_counter`, `last_pos`) readable."""
from __future__ import annotations

from .core import Wrapper


class StuckPenaltyWrapper(Wrapper):
    def __init__(self, env, max_stay=3, penalty=-0.1):
        super().__init__(env)
        self.max_stay, self.penalty = max_stay, penalty
        base = env.unwrapped
        if not hasattr(base, "enable_stuck_penalty"):
            raise TypeError("StuckPenaltyWrapper needs a CUDA-backed MERLIN env (BaseCustomEnv or BatchedMerlinEnv view)")
        base.enable_stuck_penalty(max_stay, penalty)

    @property
    def stay_counter(self):
        venv = self.unwrapped.batched
        return 0 if venv is None else int(venv.state_numpy()["stay"][0])

    @property
    def last_pos(self):
        base = self.unwrapped
        return None if base.batched is None else tuple(base.agent_pos)

    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        info.setdefault("stuck", False)
        return obs, reward, terminated, truncated, info
