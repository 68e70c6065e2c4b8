"""ORACLE (test infrastructure, NOT product code) -- CPU restatement of the code the
reference itself owns on the rollout hot path, layered on `oracle/minigrid_restated.py`.

Pinned: every function here is checked against the REAL reference modules (imported from
/root/reference over `oracle/shim.py`) by `tests/golden/make_golden.py`; the resulting
fixtures live in `tests/golden/` and `tests/test_oracle_golden.py` replays them.

What follows what (reference file:line):
  MerlinEnv.__init__           src/custom_envs/base_env.py:15-41   (max_steps = 4*size^2, see_through_walls=False)
  _layout_easy                 src/custom_envs/easy_env.py:19-39
  _layout_medium               src/custom_envs/medium_env.py:19-33
  _layout_mediumhard           src/custom_envs/medium_hard_env.py:12-45
  _layout_hard                 src/custom_envs/hard_env.py:11-73
  _layout_hardest              src/custom_envs/hardest_env.py:20-70
  _bfs_reachable               src/custom_envs/medium_hard_env.py:47-74 (== hard_env.py:75-97 == hardest_env.py:72-96)
  ThreeActions                 src/wrappers/three_action_wrapper.py:4-17
  StuckPenalty                 src/wrappers/stuck_penalty_wrapper.py:3-58
  make_env                     src/scenario_creator/scenario_creator.py:35-57 (+ src/config/scenario.yaml)
  gae_ppo                      src/ppo.py:107-120
  gae_fomaml                   src/fomaml.py:111-127 (GAE loop == src/utils/utils_rl.py:11-30)
"""
from __future__ import annotations

from collections import deque

import numpy as np

from . import minigrid_restated as mg

ENV_IDS = {
    "easy": "MERLIN-Easy-v0",
    "medium": "MERLIN-Medium-v0",
    "mediumhard": "MERLIN-MediumHard-v0",
    "hard": "MERLIN-Hard-v0",
    "hardest": "MERLIN-Hardest-v0",
}


def _bfs_reachable(grid, start, goal):
    """4-neighbour BFS over empty-or-goal cells (neighbour order: down, right, up, left)."""
    gx, gy = int(goal[0]), int(goal[1])
    sx, sy = int(start[0]), int(start[1])
    seen = {(sx, sy)}
    frontier = deque([(sx, sy)])
    while frontier:
        cx, cy = frontier.popleft()
        if (cx, cy) == (gx, gy):
            return True
        for dx, dy in ((0, 1), (1, 0), (0, -1), (-1, 0)):
            nx, ny = cx + dx, cy + dy
            if not (0 <= nx < grid.width and 0 <= ny < grid.height):
                continue
            if (nx, ny) in seen:
                continue
            cell = grid.get(nx, ny)
            if cell is None or isinstance(cell, mg.Goal) or (nx, ny) == (gx, gy):
                seen.add((nx, ny))
                frontier.append((nx, ny))
    return False


class MerlinEnv(mg.MiniGridEnv):
    """One class for the five MERLIN difficulties; `difficulty` picks the layout routine."""

    def __init__(self, difficulty="mediumhard", size=16, agent_start_pos=None, agent_start_dir=0,
                 max_steps=None, **kwargs):
        if difficulty not in ENV_IDS:
            raise ValueError(f"Unknown difficulty: {difficulty}")
        self.difficulty = difficulty
        self.size = size
        self.agent_start_pos = agent_start_pos
        self.agent_start_dir = agent_start_dir
        if max_steps is None:
            max_steps = 4 * (size ** 2)
        super().__init__(
            mission_space=mg.MissionSpace(mission_func=lambda: "reach the goal"),
            grid_size=size,
            max_steps=max_steps,
            see_through_walls=False,
            **kwargs,
        )

    # ---- layouts ---------------------------------------------------------------------
    def _gen_grid(self, width, height):
        getattr(self, "_layout_" + self.difficulty)(width, height)

    def _bordered(self, width, height):
        self.grid = mg.Grid(width, height)
        self.grid.wall_rect(0, 0, width, height)

    def _fallback_room(self, width, height):
        self._bordered(width, height)
        self.place_agent()
        self.place_obj(mg.Goal())
        self.mission = "reach the goal"

    def _layout_easy(self, width, height):
        self._bordered(width, height)
        if self.agent_start_pos is not None:
            self.agent_pos = self.agent_start_pos
            self.agent_dir = self.agent_start_dir
        else:
            self.place_agent()
        self.put_obj(mg.Goal(), width - 5, height - 5)
        self.mission = "reach the goal"

    def _layout_medium(self, width, height):
        self._bordered(width, height)
        self.place_agent()
        self.place_obj(mg.Goal())
        self.mission = "navigate to the randomly placed goal"

    def _layout_mediumhard(self, width, height):
        for _attempt in range(100):
            self._bordered(width, height)
            interior = (width - 2) * (height - 2)
            lo = max(1, int(interior * 0.10))
            hi = max(1, int(interior * 0.20)) + 1
            n_walls = self.np_random.integers(lo, hi)
            for _ in range(n_walls):
                # NB: agent_pos still holds the previous attempt's cell here (or (-1,-1))
                self.place_obj(mg.Wall(), max_tries=100)
            self.place_agent()
            goal_pos = self.place_obj(mg.Goal())
            if goal_pos is None:
                continue
            if _bfs_reachable(self.grid, self.agent_pos, goal_pos):
                self.mission = "avoid pillars and reach the goal"
                return
        print("Warning: Could not generate a valid map, returning an empty map.")
        self._fallback_room(width, height)

    def _layout_hard(self, width, height):
        self.random_goal = True
        for _attempt in range(100):
            self._bordered(width, height)
            mid = width // 2
            big = width > 10
            rows = list(range(1, height - 1))
            n_gaps = self.np_random.integers(2, 6) if big else 1
            gaps = self.np_random.choice(rows, size=n_gaps, replace=False)
            for j in range(1, height - 1):
                if j not in gaps:
                    self.grid.set(mid, j, mg.Wall())
            if big:
                n_extra = self.np_random.integers(6, 13)
                for _ in range(n_extra):
                    for _try in range(10):
                        x = self.np_random.integers(1, width - 1)
                        y = self.np_random.integers(1, height - 1)
                        if x != mid and self.grid.get(x, y) is None:
                            self.grid.set(x, y, mg.Wall())
                            break
            if not self.random_goal:
                self.put_obj(mg.Goal(), width - 2, height - 2)
                goal_pos = (width - 2, height - 2)
            else:
                goal_pos = self.place_obj(mg.Goal(), top=(mid + 1, 0), size=(width - mid - 1, height))
                if goal_pos is None:
                    continue
            if self.agent_start_pos is not None:
                self.agent_pos = self.agent_start_pos
                self.agent_dir = self.agent_start_dir
            else:
                self.place_agent(top=(1, 1), size=(mid - 1, height - 2))
            if _bfs_reachable(self.grid, self.agent_pos, goal_pos):
                self.mission = "navigate through the gaps and reach the goal"
                return
        print("Warning: Could not generate a valid map, returning an empty map.")
        self._fallback_room(width, height)

    def _layout_hardest(self, width, height):
        for _attempt in range(100):
            self._bordered(width, height)
            mx, my = width // 2, height // 2
            for y in range(1, height - 1):
                self.grid.set(mx, y, mg.Wall())
            for x in range(1, width - 1):
                self.grid.set(x, my, mg.Wall())
            self.grid.set(mx, self.np_random.integers(2, my - 1), None)
            self.grid.set(mx, self.np_random.integers(my + 1, height - 2), None)
            self.grid.set(self.np_random.integers(2, mx - 1), my, None)
            self.grid.set(self.np_random.integers(mx + 1, width - 2), my, None)
            n_obst = self.np_random.integers(6, 13)
            for _ in range(n_obst):
                x = self.np_random.integers(1, width - 1)
                y = self.np_random.integers(1, height - 1)
                if self.grid.get(x, y) is None and x != mx and y != my:
                    self.grid.set(x, y, mg.Wall())
            self.place_agent()
            goal_pos = self.place_obj(mg.Goal())
            if goal_pos is None:
                continue
            if _bfs_reachable(self.grid, self.agent_pos, goal_pos):
                self.mission = "navigate the four connected rooms to reach the goal"
                return
        self._fallback_room(width, height)


# ---- wrappers ------------------------------------------------------------------------
class ThreeActions(mg.ActionWrapper):
    """{0,1,2} -> {left, right, forward}; action_space = Discrete(3)."""

    def __init__(self, env):
        super().__init__(env)
        self.action_space = mg.Discrete(3)
        a = env.unwrapped.actions
        self._action_map = np.array([a.left, a.right, a.forward], dtype=np.int64)

    def action(self, act):
        return self._action_map[act]


class StuckPenalty(mg.Wrapper):
    """reward += penalty on every step where the position has been unchanged for >= max_stay steps."""

    def __init__(self, env, max_stay=3, penalty=-0.1):
        super().__init__(env)
        self.max_stay = max_stay
        self.penalty = penalty
        self.stay_counter = 0
        self.last_pos = None

    def reset(self, **kwargs):
        obs, info = self.env.reset(**kwargs)
        self.stay_counter = 0
        if hasattr(self.env.unwrapped, "agent_pos"):
            self.last_pos = tuple(self.env.unwrapped.agent_pos)
        return obs, info

    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        pos = tuple(self.env.unwrapped.agent_pos) if hasattr(self.env.unwrapped, "agent_pos") else None
        stuck = False
        if pos is not None and self.last_pos is not None:
            self.stay_counter = self.stay_counter + 1 if pos == self.last_pos else 0
            if self.stay_counter >= self.max_stay:
                reward += self.penalty
                stuck = True
            self.last_pos = pos
        info["stuck"] = stuck
        return obs, reward, terminated, truncated, info


def make_env(difficulty="mediumhard", size=16, fully_observable=False, flatten=False,
             stuck_penalty=False, **env_kwargs):
    """The wrapper stack `ScenarioCreator.create_env` builds (gym.make's checker wrappers are no-ops)."""
    env = MerlinEnv(difficulty=difficulty, size=size, render_mode="rgb_array", **env_kwargs)
    env = mg.FullyObsWrapper(env) if fully_observable else mg.RGBImgPartialObsWrapper(env)
    env = mg.ImgObsWrapper(env)
    if flatten:
        env = mg.FlattenObservation(env)
    env = ThreeActions(env)
    if stuck_penalty:  # the reference defines but never attaches this one
        env = StuckPenalty(env)
    return env


# ---- GAE -----------------------------------------------------------------------------
def gae_ppo(rewards, values, dones, last_value, gamma=0.99, lam=0.95):
    """torch fp32 0-dim-tensor loop, python-float gamma/lam/last_value; returns = values + adv."""
    import torch

    T = rewards.size(0)
    adv = torch.zeros_like(rewards)
    gae = 0.0
    for t in reversed(range(T)):
        mask = 1.0 - dones[t]
        next_val = last_value if t == T - 1 else values[t + 1]
        delta = rewards[t] + gamma * next_val * mask - values[t]
        gae = delta + gamma * lam * mask * gae
        adv[t] = gae
    return adv, values + adv


def gae_fomaml(rews, vals, dones, last_val, gamma=0.995, lam=0.95):
    """numpy fp32 loop; adv normalised (unbiased std) BEFORE ret = val + adv_norm."""
    import torch

    rews = np.asarray(rews, dtype=np.float32)
    vals = np.asarray(vals, dtype=np.float32)
    dones = np.asarray(dones, dtype=np.float32)
    adv = np.zeros_like(rews)
    gae = 0.0
    n = len(rews)
    for t in reversed(range(n)):
        mask = 1.0 - dones[t]
        next_v = last_val if t == n - 1 else vals[t + 1]
        delta = rews[t] + gamma * next_v * mask - vals[t]
        gae = delta + gamma * lam * mask * gae
        adv[t] = gae
    adv_t = torch.tensor(adv, dtype=torch.float32)
    adv_n = (adv_t - adv_t.mean()) / (adv_t.std() + 1e-8)
    ret_t = torch.tensor(vals) + adv_n
    return adv, adv_n.numpy(), ret_t.numpy()
