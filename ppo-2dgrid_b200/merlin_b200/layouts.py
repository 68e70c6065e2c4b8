"""Host-side layout generation for the five MERLIN difficulties, straight into packed-cell arrays.

The layouts an env sees come from `MiniGridEnv.reset()` -> `_gen_grid()`; in the B200 build they are
generated on the host (same numpy PCG64 stream, same draw order as the reference, so a given seed yields
the same grid, agent pose and goal) and uploaded to the device pool.  Reference anchors:
  easy        src/custom_envs/easy_env.py:19-39        medium   src/custom_envs/medium_env.py:19-33
  mediumhard  src/custom_envs/medium_hard_env.py:12-74  hard     src/custom_envs/hard_env.py:11-97
  hardest     src/custom_envs/hardest_env.py:20-96
and upstream minigrid 3.0.0 `place_obj` / `place_agent` / `wall_rect` (un-vendored; restated here on arrays).
"""
from __future__ import annotations

import math

import numpy as np

from . import codes

DIFFICULTIES = ("easy", "medium", "mediumhard", "hard", "hardest")
E, WALL, GOAL = codes.CODE_EMPTY, codes.CODE_WALL, codes.CODE_GOAL


class _Builder:
    """One `_gen_grid` call: a [H, W] array of packed codes plus the agent pose, fed by one RNG stream."""

    def __init__(self, rng, width, height):
        self.rng, self.W, self.H = rng, width, height
        self.agent_pos = (-1, -1)  # MiniGridEnv.reset() clears it before _gen_grid
        self.agent_dir = -1
        self.g = None

    def room(self):
        g = np.full((self.H, self.W), E, dtype=np.uint8)
        g[0, :] = g[-1, :] = WALL
        g[:, 0] = g[:, -1] = WALL
        self.g = g

    def place(self, code, top=None, size=None, max_tries=math.inf):
        """minigrid place_obj: rejection-sample an empty cell that is not the agent's; x drawn before y."""
        W, H = self.W, self.H
        top = (0, 0) if top is None else (max(top[0], 0), max(top[1], 0))
        size = (W, H) if size is None else size
        tries = 0
        while True:
            if tries > max_tries:
                raise RecursionError("rejection sampling failed in place_obj")
            tries += 1
            x = int(self.rng.integers(top[0], min(top[0] + size[0], W)))
            y = int(self.rng.integers(top[1], min(top[1] + size[1], H)))
            if self.g[y, x] != E:
                continue
            if (x, y) == self.agent_pos:
                continue
            break
        if code is not None:
            self.g[y, x] = code
        return (x, y)

    def place_agent(self, top=None, size=None):
        self.agent_pos = (-1, -1)
        self.agent_pos = self.place(None, top, size)
        self.agent_dir = int(self.rng.integers(0, 4))

    def reachable(self, goal):
        """4-neighbour flood from the agent over empty/goal cells (the three `_is_reachable` copies)."""
        free = (self.g == E) | (self.g == GOAL)
        free[goal[1], goal[0]] = True
        seen = np.zeros_like(free)
        stack = [self.agent_pos]
        seen[self.agent_pos[1], self.agent_pos[0]] = True
        while stack:
            x, y = stack.pop()
            if (x, y) == goal:
                return True
            for nx, ny in ((x, y + 1), (x + 1, y), (x, y - 1), (x - 1, y)):
                if 0 <= nx < self.W and 0 <= ny < self.H and free[ny, nx] and not seen[ny, nx]:
                    seen[ny, nx] = True
                    stack.append((nx, ny))
        return False

    def fallback(self):
        self.room()
        self.place_agent()
        self.place(GOAL)


def _easy(b):
    b.room()
    b.place_agent()
    b.g[b.H - 5, b.W - 5] = GOAL


def _medium(b):
    b.room()
    b.place_agent()
    b.place(GOAL)


def _mediumhard(b):
    for _ in range(100):
        b.room()
        interior = (b.W - 2) * (b.H - 2)
        n = int(b.rng.integers(max(1, int(interior * 0.10)), max(1, int(interior * 0.20)) + 1))
        for _ in range(n):
            b.place(WALL, max_tries=100)  # still rejects the PREVIOUS attempt's agent cell
        b.place_agent()
        goal = b.place(GOAL)
        if b.reachable(goal):
            return
    b.fallback()


def _hard(b):
    W, H = b.W, b.H
    for _ in range(100):
        b.room()
        mid = W // 2
        big = W > 10
        n_gaps = int(b.rng.integers(2, 6)) if big else 1
        gaps = b.rng.choice(list(range(1, H - 1)), size=n_gaps, replace=False)
        for j in range(1, H - 1):
            if j not in gaps:
                b.g[j, mid] = WALL
        if big:
            for _ in range(int(b.rng.integers(6, 13))):
                for _try in range(10):
                    x = int(b.rng.integers(1, W - 1))
                    y = int(b.rng.integers(1, H - 1))
                    if x != mid and b.g[y, x] == E:
                        b.g[y, x] = WALL
                        break
        goal = b.place(GOAL, top=(mid + 1, 0), size=(W - mid - 1, H))
        b.place_agent(top=(1, 1), size=(mid - 1, H - 2))
        if b.reachable(goal):
            return
    b.fallback()


def _hardest(b):
    W, H = b.W, b.H
    for _ in range(100):
        b.room()
        mx, my = W // 2, H // 2
        b.g[1:H - 1, mx] = WALL
        b.g[my, 1:W - 1] = WALL
        b.g[int(b.rng.integers(2, my - 1)), mx] = E
        b.g[int(b.rng.integers(my + 1, H - 2)), mx] = E
        b.g[my, int(b.rng.integers(2, mx - 1))] = E
        b.g[my, int(b.rng.integers(mx + 1, W - 2))] = E
        for _ in range(int(b.rng.integers(6, 13))):
            x = int(b.rng.integers(1, W - 1))
            y = int(b.rng.integers(1, H - 1))
            if b.g[y, x] == E and x != mx and y != my:
                b.g[y, x] = WALL
        b.place_agent()
        goal = b.place(GOAL)
        if b.reachable(goal):
            return
    b.fallback()


_GENERATORS = {"easy": _easy, "medium": _medium, "mediumhard": _mediumhard, "hard": _hard, "hardest": _hardest}


def generate_one(difficulty, size, rng):
    """One `reset()` worth of layout from an existing Generator (continues its stream)."""
    if difficulty not in _GENERATORS:
        raise ValueError(f"Unknown difficulty: {difficulty}")
    b = _Builder(rng, size, size)
    _GENERATORS[difficulty](b)
    return b.g.reshape(-1), np.array([b.agent_pos[0], b.agent_pos[1], b.agent_dir], dtype=np.int32)


def generate(difficulty, size, seeds):
    """Layouts of `env.reset(seed=s)` for every s in seeds -> (cells u8[L, size*size], agent i32[L, 3])."""
    seeds = list(seeds)
    cells = np.empty((len(seeds), size * size), dtype=np.uint8)
    agent = np.empty((len(seeds), 3), dtype=np.int32)
    for k, s in enumerate(seeds):
        cells[k], agent[k] = generate_one(difficulty, size, np.random.default_rng(int(s)))
    return cells, agent


def generate_stream(difficulty, size, seed, count):
    """Layouts of `reset(seed=seed)` followed by count-1 unseeded `reset()`s (one continuing RNG stream,
    as PPO training sees them: src/ppo.py:65,96)."""
    rng = np.random.default_rng(seed)
    cells = np.empty((count, size * size), dtype=np.uint8)
    agent = np.empty((count, 3), dtype=np.int32)
    for k in range(count):
        cells[k], agent[k] = generate_one(difficulty, size, rng)
    return cells, agent
