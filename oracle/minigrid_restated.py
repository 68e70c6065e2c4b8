"""ORACLE (test infrastructure, NOT product code) -- literal CPU restatement of the
`minigrid==3.0.0` / `gymnasium==1.2.1` behaviour that the reference calls into.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline legs may import
anything under `oracle/`.  The product (`ppo-2dgrid_b200/`) never does.

PARITY STATUS: **parity unpinned against upstream**.  The reference (`/root/reference`)
subclasses `minigrid.minigrid_env.MiniGridEnv` (src/custom_envs/base_env.py:5-9,35-41) and
wraps it with `minigrid.wrappers.RGBImgPartialObsWrapper` / `ImgObsWrapper`
(src/scenario_creator/scenario_creator.py:4-5,43-55), but does not vendor either package
(pins: uv.lock:467-468 minigrid 3.0.0, uv.lock:242-243 gymnasium 1.2.1) and neither is
installable here (no network).  This file restates the published upstream algorithm
per-cell / per-pixel, in the same loop order as upstream, so that it can serve as the
checker.  The reference-OWNED code that sits on top of it (the five `_gen_grid`s, the
wrappers, the GAE loops) IS pinned: `tests/golden/make_golden.py` imports the real
reference modules over `oracle/shim.py` and the fixtures it wrote are committed.

Upstream modules followed (by name, as upstream lays them out):
  minigrid/core/constants.py      -> OBJECT_TO_IDX, COLOR_TO_IDX, COLORS, STATE_TO_IDX, DIR_TO_VEC
  minigrid/core/world_object.py   -> WorldObj, Goal, Floor, Lava, Wall, Door, Key, Ball, Box
  minigrid/core/grid.py           -> Grid.{get,set,horz_wall,vert_wall,wall_rect,slice,rotate_left,
                                           process_vis,encode,render,render_tile}
  minigrid/utils/rendering.py     -> fill_coords, point_in_*, rotate_fn, highlight_img, downsample
  minigrid/minigrid_env.py        -> MiniGridEnv.{reset,step,_reward,place_obj,put_obj,place_agent,
                                           get_view_exts,gen_obs_grid,gen_obs,get_pov_render,get_frame}
  minigrid/wrappers.py            -> RGBImgPartialObsWrapper, ImgObsWrapper, FullyObsWrapper
  gymnasium/utils/seeding.py      -> np_random ; gymnasium/core.py -> Env.reset(seed=)
"""
from __future__ import annotations

import math
from enum import IntEnum

import numpy as np

# --------------------------------------------------------------------------------------
# constants (minigrid/core/constants.py)
# --------------------------------------------------------------------------------------
TILE_PIXELS = 32

COLORS = {
    "red": np.array([255, 0, 0]),
    "green": np.array([0, 255, 0]),
    "blue": np.array([0, 0, 255]),
    "purple": np.array([112, 39, 195]),
    "yellow": np.array([255, 255, 0]),
    "grey": np.array([100, 100, 100]),
}
COLOR_NAMES = sorted(list(COLORS.keys()))
COLOR_TO_IDX = {"red": 0, "green": 1, "blue": 2, "purple": 3, "yellow": 4, "grey": 5}
IDX_TO_COLOR = {v: k for k, v in COLOR_TO_IDX.items()}

OBJECT_TO_IDX = {
    "unseen": 0,
    "empty": 1,
    "wall": 2,
    "floor": 3,
    "door": 4,
    "key": 5,
    "ball": 6,
    "box": 7,
    "goal": 8,
    "lava": 9,
    "agent": 10,
}
IDX_TO_OBJECT = {v: k for k, v in OBJECT_TO_IDX.items()}

STATE_TO_IDX = {"open": 0, "closed": 1, "locked": 2}

# agent_dir -> unit step; 0 right, 1 down, 2 left, 3 up (y grows downwards)
DIR_TO_VEC = [
    np.array((1, 0)),
    np.array((0, 1)),
    np.array((-1, 0)),
    np.array((0, -1)),
]


# --------------------------------------------------------------------------------------
# rendering helpers (minigrid/utils/rendering.py)
# --------------------------------------------------------------------------------------
def downsample(img, factor):
    """Box-filter anti-aliasing: two successive float64 means (columns-in-block, then rows)."""
    assert img.shape[0] % factor == 0
    assert img.shape[1] % factor == 0
    img = img.reshape([img.shape[0] // factor, factor, img.shape[1] // factor, factor, 3])
    img = img.mean(axis=3)
    img = img.mean(axis=1)
    return img


def fill_coords(img, fn, color):
    """Paint every canvas pixel whose centre (in unit-square coordinates) satisfies fn."""
    for y in range(img.shape[0]):
        for x in range(img.shape[1]):
            yf = (y + 0.5) / img.shape[0]
            xf = (x + 0.5) / img.shape[1]
            if fn(xf, yf):
                img[y, x] = color
    return img


def rotate_fn(fin, cx, cy, theta):
    def fout(x, y):
        x = x - cx
        y = y - cy
        x2 = cx + x * math.cos(-theta) - y * math.sin(-theta)
        y2 = cy + y * math.cos(-theta) + x * math.sin(-theta)
        return fin(x2, y2)

    return fout


def point_in_line(x0, y0, x1, y1, r):
    p0 = np.array([x0, y0], dtype=np.float32)
    p1 = np.array([x1, y1], dtype=np.float32)
    dir = p1 - p0
    dist = np.linalg.norm(dir)
    dir = dir / dist

    xmin = min(x0, x1) - r
    xmax = max(x0, x1) + r
    ymin = min(y0, y1) - r
    ymax = max(y0, y1) + r

    def fn(x, y):
        if x < xmin or x > xmax or y < ymin or y > ymax:
            return False
        q = np.array([x, y])
        pq = q - p0
        a = np.dot(pq, dir)
        a = np.clip(a, 0, dist)
        p = p0 + a * dir
        dist_to_line = np.linalg.norm(q - p)
        return dist_to_line <= r

    return fn


def point_in_circle(cx, cy, r):
    def fn(x, y):
        return (x - cx) * (x - cx) + (y - cy) * (y - cy) <= r * r

    return fn


def point_in_rect(xmin, xmax, ymin, ymax):
    def fn(x, y):
        return x >= xmin and x <= xmax and y >= ymin and y <= ymax

    return fn


def point_in_triangle(a, b, c):
    a = np.array(a, dtype=np.float32)
    b = np.array(b, dtype=np.float32)
    c = np.array(c, dtype=np.float32)

    def fn(x, y):
        v0 = c - a
        v1 = b - a
        v2 = np.array((x, y)) - a

        dot00 = np.dot(v0, v0)
        dot01 = np.dot(v0, v1)
        dot02 = np.dot(v0, v2)
        dot11 = np.dot(v1, v1)
        dot12 = np.dot(v1, v2)

        inv_denom = 1 / (dot00 * dot11 - dot01 * dot01)
        u = (dot11 * dot02 - dot01 * dot12) * inv_denom
        v = (dot00 * dot12 - dot01 * dot02) * inv_denom

        return (u >= 0) and (v >= 0) and (u + v) < 1

    return fn


def highlight_img(img, color=(255, 255, 255), alpha=0.30):
    """In-place blend towards white; float64 then truncation to uint8."""
    blend_img = img + alpha * (np.array(color, dtype=np.uint8) - img)
    blend_img = blend_img.clip(0, 255).astype(np.uint8)
    img[:, :, :] = blend_img


# --------------------------------------------------------------------------------------
# world objects (minigrid/core/world_object.py)
# --------------------------------------------------------------------------------------
class WorldObj:
    def __init__(self, type, color):
        assert type in OBJECT_TO_IDX, type
        assert color in COLOR_TO_IDX, color
        self.type = type
        self.color = color
        self.contains = None
        self.init_pos = None
        self.cur_pos = None

    def can_overlap(self):
        return False

    def can_pickup(self):
        return False

    def can_contain(self):
        return False

    def see_behind(self):
        return True

    def toggle(self, env, pos):
        return False

    def encode(self):
        return (OBJECT_TO_IDX[self.type], COLOR_TO_IDX[self.color], 0)

    @staticmethod
    def decode(type_idx, color_idx, state):
        obj_type = IDX_TO_OBJECT[type_idx]
        color = IDX_TO_COLOR[color_idx]
        if obj_type == "empty" or obj_type == "unseen" or obj_type == "agent":
            return None
        is_open = state == 0
        is_locked = state == 2
        if obj_type == "wall":
            v = Wall(color)
        elif obj_type == "floor":
            v = Floor(color)
        elif obj_type == "ball":
            v = Ball(color)
        elif obj_type == "key":
            v = Key(color)
        elif obj_type == "box":
            v = Box(color)
        elif obj_type == "door":
            v = Door(color, is_open, is_locked)
        elif obj_type == "goal":
            v = Goal()
        elif obj_type == "lava":
            v = Lava()
        else:
            assert False, "unknown object type in decode '%s'" % obj_type
        return v

    def render(self, r):
        raise NotImplementedError


class Goal(WorldObj):
    def __init__(self):
        super().__init__("goal", "green")

    def can_overlap(self):
        return True

    def render(self, img):
        fill_coords(img, point_in_rect(0, 1, 0, 1), COLORS[self.color])


class Floor(WorldObj):
    def __init__(self, color="blue"):
        super().__init__("floor", color)

    def can_overlap(self):
        return True

    def render(self, img):
        color = COLORS[self.color] / 2
        fill_coords(img, point_in_rect(0.031, 1, 0.031, 1), color)


class Lava(WorldObj):
    def __init__(self):
        super().__init__("lava", "red")

    def can_overlap(self):
        return True

    def render(self, img):
        c = (255, 128, 0)
        fill_coords(img, point_in_rect(0, 1, 0, 1), c)
        for i in range(3):
            ylo = 0.3 + 0.2 * i
            yhi = 0.4 + 0.2 * i
            fill_coords(img, point_in_line(0.1, ylo, 0.3, yhi, r=0.03), (0, 0, 0))
            fill_coords(img, point_in_line(0.3, yhi, 0.5, ylo, r=0.03), (0, 0, 0))
            fill_coords(img, point_in_line(0.5, ylo, 0.7, yhi, r=0.03), (0, 0, 0))
            fill_coords(img, point_in_line(0.7, yhi, 0.9, ylo, r=0.03), (0, 0, 0))


class Wall(WorldObj):
    def __init__(self, color="grey"):
        super().__init__("wall", color)

    def see_behind(self):
        return False

    def render(self, img):
        fill_coords(img, point_in_rect(0, 1, 0, 1), COLORS[self.color])


class Door(WorldObj):
    def __init__(self, color, is_open=False, is_locked=False):
        super().__init__("door", color)
        self.is_open = is_open
        self.is_locked = is_locked

    def can_overlap(self):
        return self.is_open

    def see_behind(self):
        return self.is_open

    def toggle(self, env, pos):
        if self.is_locked:
            if isinstance(env.carrying, Key) and env.carrying.color == self.color:
                self.is_locked = False
                self.is_open = True
                return True
            return False
        self.is_open = not self.is_open
        return True

    def encode(self):
        if self.is_open:
            state = 0
        elif self.is_locked:
            state = 2
        elif not self.is_open:
            state = 1
        else:
            raise ValueError("inconsistent door state")
        return (OBJECT_TO_IDX[self.type], COLOR_TO_IDX[self.color], state)

    def render(self, img):
        c = COLORS[self.color]
        if self.is_open:
            fill_coords(img, point_in_rect(0.88, 1.00, 0.00, 1.00), c)
            fill_coords(img, point_in_rect(0.92, 0.96, 0.04, 0.96), (0, 0, 0))
            return
        if self.is_locked:
            fill_coords(img, point_in_rect(0.00, 1.00, 0.00, 1.00), c)
            fill_coords(img, point_in_rect(0.06, 0.94, 0.06, 0.94), 0.45 * np.array(c))
            fill_coords(img, point_in_rect(0.52, 0.75, 0.50, 0.56), c)
        else:
            fill_coords(img, point_in_rect(0.00, 1.00, 0.00, 1.00), c)
            fill_coords(img, point_in_rect(0.04, 0.96, 0.04, 0.96), (0, 0, 0))
            fill_coords(img, point_in_rect(0.08, 0.92, 0.08, 0.92), c)
            fill_coords(img, point_in_rect(0.12, 0.88, 0.12, 0.88), (0, 0, 0))
            fill_coords(img, point_in_circle(cx=0.75, cy=0.50, r=0.08), c)


class Key(WorldObj):
    def __init__(self, color="blue"):
        super().__init__("key", color)

    def can_pickup(self):
        return True

    def render(self, img):
        c = COLORS[self.color]
        fill_coords(img, point_in_rect(0.50, 0.63, 0.31, 0.88), c)
        fill_coords(img, point_in_rect(0.38, 0.50, 0.59, 0.66), c)
        fill_coords(img, point_in_rect(0.38, 0.50, 0.81, 0.88), c)
        fill_coords(img, point_in_circle(cx=0.56, cy=0.28, r=0.190), c)
        fill_coords(img, point_in_circle(cx=0.56, cy=0.28, r=0.064), (0, 0, 0))


class Ball(WorldObj):
    def __init__(self, color="blue"):
        super().__init__("ball", color)

    def can_pickup(self):
        return True

    def render(self, img):
        fill_coords(img, point_in_circle(0.5, 0.5, 0.31), COLORS[self.color])


class Box(WorldObj):
    def __init__(self, color, contains=None):
        super().__init__("box", color)
        self.contains = contains

    def can_pickup(self):
        return True

    def render(self, img):
        c = COLORS[self.color]
        fill_coords(img, point_in_rect(0.12, 0.88, 0.12, 0.88), c)
        fill_coords(img, point_in_rect(0.18, 0.82, 0.18, 0.82), (0, 0, 0))
        fill_coords(img, point_in_rect(0.16, 0.84, 0.47, 0.53), c)

    def toggle(self, env, pos):
        env.grid.set(pos[0], pos[1], self.contains)
        return True


# --------------------------------------------------------------------------------------
# grid (minigrid/core/grid.py)
# --------------------------------------------------------------------------------------
class Grid:
    """Row-major list of WorldObj-or-None; get(i, j) = column i (x), row j (y)."""

    tile_cache: dict = {}

    def __init__(self, width, height):
        assert width >= 3
        assert height >= 3
        self.width = width
        self.height = height
        self.grid = [None] * (width * height)

    def set(self, i, j, v):
        assert 0 <= i < self.width, f"column index {i} outside of grid of width {self.width}"
        assert 0 <= j < self.height, f"row index {j} outside of grid of height {self.height}"
        self.grid[j * self.width + i] = v

    def get(self, i, j):
        assert 0 <= i < self.width
        assert 0 <= j < self.height
        assert self.grid is not None
        return self.grid[j * self.width + i]

    def horz_wall(self, x, y, length=None, obj_type=Wall):
        if length is None:
            length = self.width - x
        for i in range(0, length):
            self.set(x + i, y, obj_type())

    def vert_wall(self, x, y, length=None, obj_type=Wall):
        if length is None:
            length = self.height - y
        for j in range(0, length):
            self.set(x, y + j, obj_type())

    def wall_rect(self, x, y, w, h):
        self.horz_wall(x, y, w)
        self.horz_wall(x, y + h - 1, w)
        self.vert_wall(x, y, h)
        self.vert_wall(x + w - 1, y, h)

    def rotate_left(self):
        """Counter-clockwise quarter turn: new(j, H'-1-i) = old(i, j)."""
        grid = Grid(self.height, self.width)
        for i in range(self.width):
            for j in range(self.height):
                v = self.get(i, j)
                grid.set(j, grid.height - 1 - i, v)
        return grid

    def slice(self, topX, topY, width, height):
        """Sub-grid; anything outside the parent becomes a (fresh) Wall."""
        grid = Grid(width, height)
        for j in range(0, height):
            for i in range(0, width):
                x = topX + i
                y = topY + j
                if 0 <= x < self.width and 0 <= y < self.height:
                    v = self.get(x, y)
                else:
                    v = Wall()
                grid.set(i, j, v)
        return grid

    @classmethod
    def render_tile(cls, obj, agent_dir=None, highlight=False, tile_size=TILE_PIXELS, subdivs=3):
        key = (agent_dir, highlight, tile_size)
        key = obj.encode() + key if obj else key
        if key in cls.tile_cache:
            return cls.tile_cache[key]

        img = np.zeros(shape=(tile_size * subdivs, tile_size * subdivs, 3), dtype=np.uint8)

        # grid lines: left column band and top row band
        fill_coords(img, point_in_rect(0, 0.031, 0, 1), (100, 100, 100))
        fill_coords(img, point_in_rect(0, 1, 0, 0.031), (100, 100, 100))

        if obj is not None:
            obj.render(img)

        if agent_dir is not None:
            tri_fn = point_in_triangle((0.12, 0.19), (0.87, 0.50), (0.12, 0.81))
            tri_fn = rotate_fn(tri_fn, cx=0.5, cy=0.5, theta=0.5 * math.pi * agent_dir)
            fill_coords(img, tri_fn, (255, 0, 0))

        if highlight:
            highlight_img(img)

        img = downsample(img, subdivs)
        cls.tile_cache[key] = img
        return img

    def render(self, tile_size, agent_pos, agent_dir=None, highlight_mask=None):
        if highlight_mask is None:
            highlight_mask = np.zeros(shape=(self.width, self.height), dtype=bool)

        width_px = self.width * tile_size
        height_px = self.height * tile_size
        img = np.zeros(shape=(height_px, width_px, 3), dtype=np.uint8)

        for j in range(0, self.height):
            for i in range(0, self.width):
                cell = self.get(i, j)
                agent_here = np.array_equal(agent_pos, (i, j))
                tile_img = Grid.render_tile(
                    cell,
                    agent_dir=agent_dir if agent_here else None,
                    highlight=highlight_mask[i, j],
                    tile_size=tile_size,
                )
                ymin = j * tile_size
                ymax = (j + 1) * tile_size
                xmin = i * tile_size
                xmax = (i + 1) * tile_size
                img[ymin:ymax, xmin:xmax, :] = tile_img  # float64 -> uint8 truncation
        return img

    def encode(self, vis_mask=None):
        if vis_mask is None:
            vis_mask = np.ones((self.width, self.height), dtype=bool)
        array = np.zeros((self.width, self.height, 3), dtype="uint8")
        for i in range(self.width):
            for j in range(self.height):
                if vis_mask[i, j]:
                    v = self.get(i, j)
                    if v is None:
                        array[i, j, 0] = OBJECT_TO_IDX["empty"]
                        array[i, j, 1] = 0
                        array[i, j, 2] = 0
                    else:
                        array[i, j, :] = v.encode()
        return array

    @staticmethod
    def decode(array):
        width, height, channels = array.shape
        assert channels == 3
        vis_mask = np.ones(shape=(width, height), dtype=bool)
        grid = Grid(width, height)
        for i in range(width):
            for j in range(height):
                type_idx, color_idx, state = array[i, j]
                v = WorldObj.decode(int(type_idx), int(color_idx), int(state))
                grid.set(i, j, v)
                vis_mask[i, j] = type_idx != OBJECT_TO_IDX["unseen"]
        return grid, vis_mask

    def process_vis(self, agent_pos):
        """Shadow-casting-ish visibility sweep, bottom row upwards; erases unseen cells."""
        mask = np.zeros(shape=(self.width, self.height), dtype=bool)
        mask[agent_pos[0], agent_pos[1]] = True

        for j in reversed(range(0, self.height)):
            for i in range(0, self.width - 1):
                if not mask[i, j]:
                    continue
                cell = self.get(i, j)
                if cell and not cell.see_behind():
                    continue
                mask[i + 1, j] = True
                if j > 0:
                    mask[i + 1, j - 1] = True
                    mask[i, j - 1] = True

            for i in reversed(range(1, self.width)):
                if not mask[i, j]:
                    continue
                cell = self.get(i, j)
                if cell and not cell.see_behind():
                    continue
                mask[i - 1, j] = True
                if j > 0:
                    mask[i - 1, j - 1] = True
                    mask[i, j - 1] = True

        for j in range(0, self.height):
            for i in range(0, self.width):
                if not mask[i, j]:
                    self.set(i, j, None)

        return mask


# --------------------------------------------------------------------------------------
# minimal gymnasium pieces the path touches (gymnasium/core.py, spaces, seeding)
# --------------------------------------------------------------------------------------
class Discrete:
    def __init__(self, n):
        self.n = int(n)
        self._rng = np.random.default_rng()

    def sample(self):
        return int(self._rng.integers(0, self.n))

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)

    def contains(self, x):
        return 0 <= int(x) < self.n

    def __repr__(self):
        return f"Discrete({self.n})"


class BoxSpace:
    def __init__(self, low, high, shape, dtype):
        self.low, self.high, self.shape, self.dtype = low, high, tuple(shape), np.dtype(dtype)


class DictSpace:
    def __init__(self, spaces):
        self.spaces = dict(spaces)


def np_random(seed=None):
    """gymnasium.utils.seeding.np_random: Generator(PCG64(SeedSequence(seed)))."""
    seed_seq = np.random.SeedSequence(seed)
    np_seed = seed_seq.entropy
    rng = np.random.Generator(np.random.PCG64(seed_seq))
    return rng, np_seed


class Env:
    """gymnasium.Env: lazily seeded `np_random`; `reset(seed=)` reseeds only if seed given."""

    _np_random = None
    _np_random_seed = None
    render_mode = None

    @property
    def unwrapped(self):
        return self

    @property
    def np_random(self):
        if self._np_random is None:
            self._np_random, self._np_random_seed = np_random()
        return self._np_random

    @np_random.setter
    def np_random(self, value):
        self._np_random = value
        self._np_random_seed = -1

    def reset(self, *, seed=None, options=None):
        if seed is not None:
            self._np_random, self._np_random_seed = np_random(seed)

    def close(self):
        pass


class Wrapper:
    """gymnasium.Wrapper: forwards everything to `env`."""

    def __init__(self, env):
        self.env = env
        self._action_space = None
        self._observation_space = None

    def __getattr__(self, name):
        if name.startswith("_"):
            raise AttributeError(name)
        return getattr(self.env, name)

    @property
    def unwrapped(self):
        return self.env.unwrapped

    @property
    def action_space(self):
        if self._action_space is None:
            return self.env.action_space
        return self._action_space

    @action_space.setter
    def action_space(self, space):
        self._action_space = space

    @property
    def observation_space(self):
        if self._observation_space is None:
            return self.env.observation_space
        return self._observation_space

    @observation_space.setter
    def observation_space(self, space):
        self._observation_space = space

    def reset(self, *, seed=None, options=None):
        return self.env.reset(seed=seed, options=options)

    def step(self, action):
        return self.env.step(action)

    def close(self):
        return self.env.close()


class ObservationWrapper(Wrapper):
    def reset(self, *, seed=None, options=None):
        obs, info = self.env.reset(seed=seed, options=options)
        return self.observation(obs), info

    def step(self, action):
        obs, reward, terminated, truncated, info = self.env.step(action)
        return self.observation(obs), reward, terminated, truncated, info

    def observation(self, obs):
        raise NotImplementedError


class ActionWrapper(Wrapper):
    def step(self, action):
        return self.env.step(self.action(action))

    def action(self, action):
        raise NotImplementedError


class FlattenObservation(ObservationWrapper):
    """gymnasium.wrappers.FlattenObservation for a Box image: C-order ravel."""

    def __init__(self, env):
        super().__init__(env)
        shp = env.observation_space.shape
        self.observation_space = BoxSpace(0, 255, (int(np.prod(shp)),), env.observation_space.dtype)

    def observation(self, obs):
        return np.asarray(obs).reshape(-1)


# --------------------------------------------------------------------------------------
# the environment (minigrid/minigrid_env.py)
# --------------------------------------------------------------------------------------
class Actions(IntEnum):
    left = 0
    right = 1
    forward = 2
    pickup = 3
    drop = 4
    toggle = 5
    done = 6


class MissionSpace:
    """minigrid.core.mission.MissionSpace -- only what BaseCustomEnv needs."""

    def __init__(self, mission_func, ordered_placeholders=None, seed=None):
        self.mission_func = mission_func
        self.ordered_placeholders = ordered_placeholders

    def sample(self):
        return self.mission_func()


class MiniGridEnv(Env):
    Actions = Actions

    def __init__(
        self,
        mission_space,
        grid_size=None,
        width=None,
        height=None,
        max_steps=100,
        see_through_walls=False,
        agent_view_size=7,
        render_mode=None,
        screen_size=640,
        highlight=True,
        tile_size=TILE_PIXELS,
        agent_pov=False,
    ):
        self.mission = mission_space.sample()
        if grid_size:
            assert width is None and height is None
            width = grid_size
            height = grid_size
        assert width is not None and height is not None

        self.actions = Actions
        self.action_space = Discrete(len(self.actions))

        assert agent_view_size % 2 == 1
        assert agent_view_size >= 3
        self.agent_view_size = agent_view_size

        image_space = BoxSpace(0, 255, (self.agent_view_size, self.agent_view_size, 3), "uint8")
        self.observation_space = DictSpace(
            {"image": image_space, "direction": Discrete(4), "mission": mission_space}
        )
        self.reward_range = (0, 1)

        self.screen_size = screen_size
        self.render_size = None
        self.window = None
        self.clock = None

        self.width = width
        self.height = height
        assert isinstance(max_steps, int), f"max_steps must be int, got {type(max_steps)}"
        self.max_steps = max_steps
        self.see_through_walls = see_through_walls

        self.agent_pos = None
        self.agent_dir = None
        self.grid = Grid(width, height)
        self.carrying = None

        self.render_mode = render_mode
        self.highlight = highlight
        self.tile_size = tile_size
        self.agent_pov = agent_pov

    # -- episode control ---------------------------------------------------------------
    def reset(self, *, seed=None, options=None):
        super().reset(seed=seed)
        self.agent_pos = (-1, -1)
        self.agent_dir = -1
        self._gen_grid(self.width, self.height)

        assert (
            self.agent_pos >= (0, 0)
            if isinstance(self.agent_pos, tuple)
            else all(self.agent_pos >= 0) and self.agent_dir >= 0
        )
        start_cell = self.grid.get(*self.agent_pos)
        assert start_cell is None or start_cell.can_overlap()

        self.carrying = None
        self.step_count = 0
        obs = self.gen_obs()
        return obs, {}

    @property
    def steps_remaining(self):
        return self.max_steps - self.step_count

    def _gen_grid(self, width, height):
        raise NotImplementedError

    def _reward(self):
        return 1 - 0.9 * (self.step_count / self.max_steps)

    def _rand_int(self, low, high):
        return self.np_random.integers(low, high)

    # -- placement ---------------------------------------------------------------------
    def place_obj(self, obj, top=None, size=None, reject_fn=None, max_tries=math.inf):
        if top is None:
            top = (0, 0)
        else:
            top = (max(top[0], 0), max(top[1], 0))
        if size is None:
            size = (self.grid.width, self.grid.height)

        num_tries = 0
        while True:
            if num_tries > max_tries:
                raise RecursionError("rejection sampling failed in place_obj")
            num_tries += 1
            pos = (
                self._rand_int(top[0], min(top[0] + size[0], self.grid.width)),
                self._rand_int(top[1], min(top[1] + size[1], self.grid.height)),
            )
            if self.grid.get(*pos) is not None:
                continue
            if np.array_equal(pos, self.agent_pos):
                continue
            if reject_fn and reject_fn(self, pos):
                continue
            break

        self.grid.set(pos[0], pos[1], obj)
        if obj is not None:
            obj.init_pos = pos
            obj.cur_pos = pos
        return pos

    def put_obj(self, obj, i, j):
        self.grid.set(i, j, obj)
        obj.init_pos = (i, j)
        obj.cur_pos = (i, j)

    def place_agent(self, top=None, size=None, rand_dir=True, max_tries=math.inf):
        self.agent_pos = (-1, -1)
        pos = self.place_obj(None, top, size, max_tries=max_tries)
        self.agent_pos = pos
        if rand_dir:
            self.agent_dir = self._rand_int(0, 4)
        return pos

    # -- geometry ----------------------------------------------------------------------
    @property
    def dir_vec(self):
        assert 0 <= self.agent_dir < 4, f"Invalid agent_dir: {self.agent_dir}"
        return DIR_TO_VEC[self.agent_dir]

    @property
    def right_vec(self):
        dx, dy = self.dir_vec
        return np.array((-dy, dx))

    @property
    def front_pos(self):
        return self.agent_pos + self.dir_vec

    def get_view_exts(self, agent_view_size=None):
        agent_view_size = agent_view_size or self.agent_view_size
        if self.agent_dir == 0:  # facing right
            topX = self.agent_pos[0]
            topY = self.agent_pos[1] - agent_view_size // 2
        elif self.agent_dir == 1:  # facing down
            topX = self.agent_pos[0] - agent_view_size // 2
            topY = self.agent_pos[1]
        elif self.agent_dir == 2:  # facing left
            topX = self.agent_pos[0] - agent_view_size + 1
            topY = self.agent_pos[1] - agent_view_size // 2
        elif self.agent_dir == 3:  # facing up
            topX = self.agent_pos[0] - agent_view_size // 2
            topY = self.agent_pos[1] - agent_view_size + 1
        else:
            assert False, "invalid agent direction"
        botX = topX + agent_view_size
        botY = topY + agent_view_size
        return topX, topY, botX, botY

    # -- dynamics ----------------------------------------------------------------------
    def step(self, action):
        self.step_count += 1
        reward = 0
        terminated = False
        truncated = False

        fwd_pos = self.front_pos
        fwd_cell = self.grid.get(*fwd_pos)

        if action == self.actions.left:
            self.agent_dir -= 1
            if self.agent_dir < 0:
                self.agent_dir += 4
        elif action == self.actions.right:
            self.agent_dir = (self.agent_dir + 1) % 4
        elif action == self.actions.forward:
            if fwd_cell is None or fwd_cell.can_overlap():
                self.agent_pos = tuple(fwd_pos)
            if fwd_cell is not None and fwd_cell.type == "goal":
                terminated = True
                reward = self._reward()
            if fwd_cell is not None and fwd_cell.type == "lava":
                terminated = True
        elif action == self.actions.pickup:
            if fwd_cell and fwd_cell.can_pickup():
                if self.carrying is None:
                    self.carrying = fwd_cell
                    self.carrying.cur_pos = np.array([-1, -1])
                    self.grid.set(fwd_pos[0], fwd_pos[1], None)
        elif action == self.actions.drop:
            if not fwd_cell and self.carrying:
                self.grid.set(fwd_pos[0], fwd_pos[1], self.carrying)
                self.carrying.cur_pos = fwd_pos
                self.carrying = None
        elif action == self.actions.toggle:
            if fwd_cell:
                fwd_cell.toggle(self, fwd_pos)
        elif action == self.actions.done:
            pass
        else:
            raise ValueError(f"Unknown action: {action}")

        if self.step_count >= self.max_steps:
            truncated = True

        obs = self.gen_obs()
        return obs, reward, terminated, truncated, {}

    # -- observation -------------------------------------------------------------------
    def gen_obs_grid(self, agent_view_size=None):
        topX, topY, botX, botY = self.get_view_exts(agent_view_size)
        agent_view_size = agent_view_size or self.agent_view_size

        grid = self.grid.slice(topX, topY, agent_view_size, agent_view_size)
        for i in range(self.agent_dir + 1):
            grid = grid.rotate_left()

        if not self.see_through_walls:
            vis_mask = grid.process_vis(agent_pos=(agent_view_size // 2, agent_view_size - 1))
        else:
            vis_mask = np.ones(shape=(grid.width, grid.height), dtype=bool)

        agent_pos = grid.width // 2, grid.height - 1
        if self.carrying:
            grid.set(*agent_pos, self.carrying)
        else:
            grid.set(*agent_pos, None)
        return grid, vis_mask

    def gen_obs(self):
        grid, vis_mask = self.gen_obs_grid()
        image = grid.encode(vis_mask)
        obs = {"image": image, "direction": self.agent_dir, "mission": self.mission}
        return obs

    def get_pov_render(self, tile_size):
        grid, vis_mask = self.gen_obs_grid()
        img = grid.render(
            tile_size,
            agent_pos=(self.agent_view_size // 2, self.agent_view_size - 1),
            agent_dir=3,
            highlight_mask=vis_mask,
        )
        return img

    def get_full_render(self, highlight, tile_size):
        _, vis_mask = self.gen_obs_grid()
        f_vec = self.dir_vec
        r_vec = self.right_vec
        top_left = (
            self.agent_pos
            + f_vec * (self.agent_view_size - 1)
            - r_vec * (self.agent_view_size // 2)
        )
        highlight_mask = np.zeros(shape=(self.width, self.height), dtype=bool)
        for vis_j in range(0, self.agent_view_size):
            for vis_i in range(0, self.agent_view_size):
                if not vis_mask[vis_i, vis_j]:
                    continue
                abs_i, abs_j = top_left - (f_vec * vis_j) + (r_vec * vis_i)
                if abs_i < 0 or abs_i >= self.width:
                    continue
                if abs_j < 0 or abs_j >= self.height:
                    continue
                highlight_mask[abs_i, abs_j] = True
        img = self.grid.render(
            tile_size,
            self.agent_pos,
            self.agent_dir,
            highlight_mask=highlight_mask if highlight else None,
        )
        return img

    def get_frame(self, highlight=True, tile_size=TILE_PIXELS, agent_pov=False):
        if agent_pov:
            return self.get_pov_render(tile_size)
        return self.get_full_render(highlight, tile_size)

    def render(self):
        img = self.get_frame(self.highlight, self.tile_size, self.agent_pov)
        if self.render_mode == "rgb_array":
            return img
        return None


# --------------------------------------------------------------------------------------
# observation wrappers (minigrid/wrappers.py)
# --------------------------------------------------------------------------------------
class FullyObsWrapper(ObservationWrapper):
    def __init__(self, env):
        super().__init__(env)
        u = self.env.unwrapped
        new_image_space = BoxSpace(0, 255, (u.width, u.height, 3), "uint8")
        self.observation_space = DictSpace({**self.env.observation_space.spaces, "image": new_image_space})

    def observation(self, obs):
        env = self.unwrapped
        full_grid = env.grid.encode()
        full_grid[env.agent_pos[0]][env.agent_pos[1]] = np.array(
            [OBJECT_TO_IDX["agent"], COLOR_TO_IDX["red"], env.agent_dir]
        )
        return {**obs, "image": full_grid}


class RGBImgPartialObsWrapper(ObservationWrapper):
    def __init__(self, env, tile_size=8):
        super().__init__(env)
        self.tile_size = tile_size
        obs_shape = env.observation_space.spaces["image"].shape
        new_image_space = BoxSpace(
            0, 255, (obs_shape[0] * tile_size, obs_shape[1] * tile_size, 3), "uint8"
        )
        self.observation_space = DictSpace({**self.env.observation_space.spaces, "image": new_image_space})

    def observation(self, obs):
        rgb_img_partial = self.unwrapped.get_frame(tile_size=self.tile_size, agent_pov=True)
        return {**obs, "image": rgb_img_partial}


class ImgObsWrapper(ObservationWrapper):
    def __init__(self, env):
        super().__init__(env)
        self.observation_space = env.observation_space.spaces["image"]

    def observation(self, obs):
        return obs["image"]
