"""BaseCustomEnv -- one MERLIN / MiniGrid environment with the reference's single-env interface, executed by
the batched CUDA kernels at N = 1.

Mirror of reference src/custom_envs/base_env.py:9-53 (constants: grid_size = size, max_steps = 4 * size^2,
see_through_walls = False, agent_view_size 7) plus the parts of upstream `MiniGridEnv` the reference's callers
touch: `reset(seed=)`, `step(a)` with the 7-action set, `.agent_pos/.agent_dir/.step_count/.grid/.actions`,
`get_frame()`, `close()`.  Layout generation (`_gen_grid`) runs on the host from the env's numpy Generator
(merlin_b200.layouts follows the reference's draw order); step / gen_obs / process_vis / encode / RGB POV
rendering run on the GPU through libmerlin_b200.so.  There is no CPU path: the first `reset()` needs CUDA.
"""
from __future__ import annotations

from enum import IntEnum

import numpy as np

from merlin_b200 import codes, layouts, tiles

from ..spaces import Box, Discrete


class Actions(IntEnum):
    left = 0
    right = 1
    forward = 2
    pickup = 3
    drop = 4
    toggle = 5
    done = 6


class GridView:
    """Read-only snapshot of an env's grid with the `Grid` accessors callers use (`get`, `encode`, size)."""

    _NAMES = {codes.WALL: "wall", codes.FLOOR: "floor", codes.DOOR_OPEN: "door", codes.KEY: "key", codes.BALL: "ball",
              codes.BOX: "box", codes.GOAL: "goal", codes.LAVA: "lava", codes.DOOR_CLOSED: "door",
              codes.DOOR_LOCKED: "door"}

    class Cell:
        def __init__(self, code):
            self.code = int(code)
            self.type = GridView._NAMES[self.code & 0xF]
            self.color = codes.COLOR_NAMES[(self.code >> 4) & 7]

        def can_overlap(self):
            return (self.code & 0xF) in (codes.FLOOR, codes.DOOR_OPEN, codes.GOAL, codes.LAVA)

        def see_behind(self):
            return (self.code & 0xF) not in (codes.WALL, codes.DOOR_CLOSED, codes.DOOR_LOCKED)

    def __init__(self, cells, width, height):
        self.width, self.height = width, height
        self.cells = np.asarray(cells, dtype=np.uint8)[: width * height].reshape(height, width)

    def get(self, i, j):
        code = self.cells[j, i]
        return None if (code & 0xF) == codes.EMPTY else GridView.Cell(code)

    def encode(self):
        return codes.unpack_to_encoding(self.cells.reshape(1, -1), self.width, self.height)[0]


class BaseCustomEnv:
    difficulty = None  # set by the five subclasses
    metadata = {"render_modes": ["human", "rgb_array"], "render_fps": 10}

    def __init__(self, size=16, agent_start_pos=None, agent_start_dir=0, max_steps=None, render_mode=None,
                 device=None, tile_size=8, **kwargs):
        if self.difficulty is None:
            raise NotImplementedError("Subclasses must set `difficulty` (the layout routine) -- see register.py")
        self.size = self.width = self.height = int(size)
        self.agent_start_pos, self.agent_start_dir = agent_start_pos, agent_start_dir
        self.max_steps = int(max_steps) if max_steps is not None else 4 * self.size ** 2
        self.see_through_walls = False
        self.agent_view_size = 7
        self.tile_size = tile_size
        self.render_mode = render_mode
        self.actions = Actions
        self.action_space = Discrete(len(Actions))
        self.observation_space = {"image": Box(0, 255, (7, 7, 3), np.uint8), "direction": Discrete(4)}
        self.mission = "reach the goal"
        self.device = device
        self.np_random = None
        self._venv = None
        self._stuck_cfg = None
        self._obs_rgb = None
        self._state = None
        self.unwrapped = self

    # ---- device side ---------------------------------------------------------------------------------------
    def _make_venv(self):
        import torch
        from merlin_b200 import BatchedMerlinEnv

        dev = self.device if self.device is not None else torch.device("cuda", torch.cuda.current_device() if torch.cuda.is_available() else 0)
        cells = np.full((1, self.width * self.height), codes.CODE_EMPTY, dtype=np.uint8)
        kw = {}
        if self._stuck_cfg is not None:
            kw = dict(stuck_penalty=True, stuck_max_stay=self._stuck_cfg[0], stuck_penalty_value=self._stuck_cfg[1])
        self._venv = BatchedMerlinEnv(1, cells, np.array([[1, 1, 0]], dtype=np.int32), width=self.width,
                                      height=self.height, max_steps=self.max_steps, device=dev, n_actions=7,
                                      auto_reset=False, **kw)
        self._act = torch.zeros(1, dtype=torch.int64, device=self._venv.device)

    def enable_stuck_penalty(self, max_stay=3, penalty=-0.1):
        """Used by StuckPenaltyWrapper: the counter and the penalty are applied inside the step kernel."""
        self._stuck_cfg = (int(max_stay), float(penalty))
        if self._venv is not None:
            self._venv.close()
            self._venv = None

    # ---- gymnasium-style interface -------------------------------------------------------------------------
    def _gen_layout(self):
        if self.difficulty == "easy" and self.agent_start_pos is not None:
            cells, _ = layouts.generate_one("easy", self.size, np.random.default_rng(0))
            return cells, np.array([self.agent_start_pos[0], self.agent_start_pos[1], self.agent_start_dir], dtype=np.int32)
        return layouts.generate_one(self.difficulty, self.size, self.np_random)

    def _obs(self, sym):
        return {"image": sym, "direction": int(self.agent_dir), "mission": self.mission}

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self.np_random = np.random.default_rng(seed)
        elif self.np_random is None:
            self.np_random = np.random.default_rng()  # lazily OS-seeded, like gymnasium
        if self._venv is None:
            self._make_venv()
        cells, agent = self._gen_layout()
        self._venv.upload_layouts(cells.reshape(1, -1), agent.reshape(1, 3))
        self._venv.set_cursors(np.zeros(1, dtype=np.int32))
        rgb, sym = self._venv.reset()
        self._obs_rgb = rgb[0].cpu().numpy()
        self._state = None
        return self._obs(sym[0].cpu().numpy()), {}

    def step(self, action):
        if self._venv is None:
            raise RuntimeError("Cannot call env.step() before calling env.reset()")
        a = int(action)
        if not 0 <= a < len(Actions):
            raise ValueError(f"Unknown action: {action}")
        self._act.fill_(a)
        rgb, rew, term, trunc, info = self._venv.step(self._act)
        self._obs_rgb = rgb[0].cpu().numpy()
        self._state = None
        sym = info["obs_symbolic"][0].cpu().numpy()
        out_info = {"stuck": bool(info["stuck"][0].item())} if self._stuck_cfg is not None else {}
        return self._obs(sym), float(rew[0].item()), bool(term[0].item()), bool(trunc[0].item()), out_info

    # ---- state views ---------------------------------------------------------------------------------------
    def _s(self):
        if self._state is None:
            self._state = self._venv.state_numpy()
        return self._state

    @property
    def agent_pos(self):
        if self._venv is None:
            return (-1, -1)
        s = self._s()
        return (int(s["x"][0]), int(s["y"][0]))

    @property
    def agent_dir(self):
        return -1 if self._venv is None else int(self._s()["dir"][0])

    @property
    def step_count(self):
        return 0 if self._venv is None else int(self._s()["step_count"][0])

    @property
    def grid(self):
        return GridView(self._venv.cells_numpy()[0], self.width, self.height)

    @property
    def batched(self):
        """The underlying N = 1 BatchedMerlinEnv (device tensors)."""
        return self._venv

    def gen_full_obs(self):
        """FullyObsWrapper's image for the current state: u8[W, H, 3], computed on the device."""
        return self._venv.full_observation()[0].cpu().numpy()

    def get_pov_render(self, tile_size=None):
        """The 56x56x3 egocentric frame of the current state (RGBImgPartialObsWrapper's observation)."""
        return self._obs_rgb

    def get_full_render(self, highlight=True, tile_size=32):
        """Whole-grid frame for human viewing (upstream MiniGridEnv.get_full_render): host-side tile blit, not on
        the training path.  `highlight` marks the cells the agent currently sees."""
        W, H = self.width, self.height
        cells = self._venv.cells_numpy()[0][: W * H].reshape(H, W)
        ax, ay = self.agent_pos
        d = self.agent_dir
        vis = np.zeros((H, W), dtype=bool)
        if highlight:
            vis = _visible_world_mask(cells, ax, ay, d)
        img = np.zeros((H * tile_size, W * tile_size, 3), dtype=np.uint8)
        cache = {}
        for j in range(H):
            for i in range(W):
                code = int(cells[j, i])
                key = (code, d if (i, j) == (ax, ay) else None, bool(vis[j, i]))
                if key not in cache:
                    cache[key] = tiles.render_tile(code & 0xF, (code >> 4) & 7, agent=key[1] is not None, highlight=key[2],
                                                   tile=tile_size, agent_dir=key[1] if key[1] is not None else 3)
                img[j * tile_size:(j + 1) * tile_size, i * tile_size:(i + 1) * tile_size] = cache[key]
        return img

    def get_frame(self, highlight=True, tile_size=32, agent_pov=False):
        return self.get_pov_render(tile_size) if agent_pov else self.get_full_render(highlight, tile_size)

    def render(self):
        return self.get_frame(True, 32, False) if self.render_mode == "rgb_array" else None

    def close(self):
        if self._venv is not None:
            self._venv.close()
            self._venv = None


def _visible_world_mask(cells, ax, ay, d):
    """World-space visibility mask of the 7x7 view (host helper for get_full_render only)."""
    H, W = cells.shape
    fx, fy = ((1, 0), (0, 1), (-1, 0), (0, -1))[d]
    rx, ry = -fy, fx
    opaque = (codes.WALL, codes.DOOR_CLOSED, codes.DOOR_LOCKED)
    world = {}
    transp = np.zeros((7, 7), dtype=bool)  # [vi][vj]
    for vi in range(7):
        for vj in range(7):
            x, y = ax + (6 - vj) * fx + (vi - 3) * rx, ay + (6 - vj) * fy + (vi - 3) * ry
            inb = 0 <= x < W and 0 <= y < H
            world[(vi, vj)] = (x, y) if inb else None
            transp[vi, vj] = inb and (int(cells[y, x]) & 0xF) not in opaque
    mask = np.zeros((7, 7), dtype=bool)
    mask[3, 6] = True
    for j in range(6, -1, -1):
        for i in range(0, 6):
            if mask[i, j] and transp[i, j]:
                mask[i + 1, j] = True
                if j > 0:
                    mask[i + 1, j - 1] = True
                    mask[i, j - 1] = True
        for i in range(6, 0, -1):
            if mask[i, j] and transp[i, j]:
                mask[i - 1, j] = True
                if j > 0:
                    mask[i - 1, j - 1] = True
                    mask[i, j - 1] = True
    out = np.zeros((H, W), dtype=bool)
    for (vi, vj), xy in world.items():
        if xy is not None and mask[vi, vj]:
            out[xy[1], xy[0]] = True
    return out
