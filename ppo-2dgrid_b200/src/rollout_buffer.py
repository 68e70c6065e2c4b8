"""RolloutBuffer -- pre-allocated rollout storage (reference src/rollout_buffer.py:3-32: same constructor,
`add`, `get`, `ptr`, `max_size` and tensor names).

Batched, device-resident form: with `num_envs = N > 1` every array is time-major `[T, N, ...]`
(T = buffer_size // N), which is the layout the GAE kernel scans (coalesced over envs, sequential in t) and the
layout the env kernel writes frames into: `obs_slot(t)` hands out the `[N, 56, 56, 3]` slice of `states` so
`step(..., out_obs=...)` renders the next observation straight into the rollout, no copy.  Frames are kept as
uint8 when `obs_dtype=torch.uint8` (the policy casts on read): 9 408 B per step instead of the reference's
37 632 B of float32.
"""
from __future__ import annotations

import torch


class RolloutBuffer:
    def __init__(self, buffer_size, obs_shape, device, is_discrete=True, num_envs=1, obs_dtype=torch.float32):
        self.num_envs = int(num_envs)
        if buffer_size % self.num_envs:
            raise ValueError(f"buffer_size {buffer_size} is not a multiple of num_envs {num_envs}")
        self.horizon = buffer_size // self.num_envs
        lead = (self.horizon,) if self.num_envs == 1 else (self.horizon, self.num_envs)
        self.states = torch.zeros(lead + tuple(obs_shape), dtype=obs_dtype, device=device)
        self.actions = torch.zeros(lead, dtype=torch.long if is_discrete else torch.float32, device=device)
        self.logprobs = torch.zeros(lead, dtype=torch.float32, device=device)
        self.rewards = torch.zeros(lead, dtype=torch.float32, device=device)
        self.values = torch.zeros(lead, dtype=torch.float32, device=device)
        self.dones = torch.zeros(lead, dtype=torch.float32, device=device)
        self.max_size = self.horizon
        self.ptr = 0

    def obs_slot(self, t):
        """View of `states[t]` as `[N, *obs_shape]` -- a valid `out_obs` target for BatchedMerlinEnv."""
        s = self.states[t]
        return s.unsqueeze(0) if self.num_envs == 1 else s

    def add(self, state, action, logprob, value, reward, done):
        """One time step: scalars / `[*obs_shape]` for a single env, `[N]` / `[N, *obs_shape]` tensors when batched.
        `state=None` means the frame is already in place (written through `obs_slot`)."""
        p = self.ptr
        if state is not None:
            self.states[p] = state
        self.actions[p] = action
        self.logprobs[p] = logprob
        self.values[p] = value
        self.rewards[p] = reward
        self.dones[p] = done
        self.ptr = (p + 1) % self.max_size

    def get(self):
        self.ptr = 0
        return self.states, self.actions, self.logprobs, self.rewards, self.values, self.dones
