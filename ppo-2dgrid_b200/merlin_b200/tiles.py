"""Tile atlas for the RGB observation: the 8x8x3 tiles `Grid.render_tile` (minigrid 3.0.0) would produce
for every packed cell code, rendered on the host with array operations and uploaded once.

Replaces the tile cache that `RGBImgPartialObsWrapper.observation` -> `get_frame(tile_size=8,
agent_pov=True)` -> `Grid.render` fills lazily (reference call site:
src/scenario_creator/scenario_creator.py:48).  Because invisible cells are erased before rendering and
the highlight mask equals the visibility mask, a POV frame only ever contains: the un-highlighted empty
tile (invisible cell, atlas slot 0), highlighted object tiles (slot = packed code) and the agent cell
(slot 10 = red triangle over empty; 13/14/15 | colour<<4 = triangle over a carried key/ball/box).

Pipeline per tile (all as upstream): 24x24 u8 canvas (3x supersampling) -> grid lines -> object ->
agent triangle (dir 3) -> highlight (float64 blend, truncation) -> 3x3 box mean (float64, two passes)
-> truncation to u8.
"""
from __future__ import annotations

import math

import numpy as np

from . import codes

TILE = 8
SUBDIVS = 3
N_TILES = 128

_COLORS = np.array([[255, 0, 0], [0, 255, 0], [0, 0, 255], [112, 39, 195], [255, 255, 0], [100, 100, 100]])


def _sample_grid(n):
    c = (np.arange(n) + 0.5) / n
    return np.meshgrid(c, c, indexing="xy")  # x varies along columns, y along rows


def _rect(x, y, xmin, xmax, ymin, ymax):
    return (x >= xmin) & (x <= xmax) & (y >= ymin) & (y <= ymax)


def _circle(x, y, cx, cy, r):
    return (x - cx) * (x - cx) + (y - cy) * (y - cy) <= r * r


def _line(x, y, x0, y0, x1, y1, r):
    p0 = np.array([x0, y0], dtype=np.float32)
    p1 = np.array([x1, y1], dtype=np.float32)
    d = p1 - p0
    dist = np.linalg.norm(d)
    d = d / dist
    inside_box = ~((x < min(x0, x1) - r) | (x > max(x0, x1) + r) | (y < min(y0, y1) - r) | (y > max(y0, y1) + r))
    pqx, pqy = x - p0[0], y - p0[1]
    a = np.clip(pqx * d[0] + pqy * d[1], 0, dist)
    px, py = p0[0] + a * d[0], p0[1] + a * d[1]
    return inside_box & (np.sqrt((x - px) ** 2 + (y - py) ** 2) <= r)


def _triangle(x, y, a, b, c):
    a = np.array(a, dtype=np.float32)
    b = np.array(b, dtype=np.float32)
    c = np.array(c, dtype=np.float32)
    v0, v1 = c - a, b - a
    v2x, v2y = x - a[0], y - a[1]
    dot00, dot01, dot11 = np.dot(v0, v0), np.dot(v0, v1), np.dot(v1, v1)
    dot02 = v0[0] * v2x + v0[1] * v2y
    dot12 = v1[0] * v2x + v1[1] * v2y
    inv = 1 / (dot00 * dot11 - dot01 * dot01)
    u = (dot11 * dot02 - dot01 * dot12) * inv
    v = (dot00 * dot12 - dot01 * dot02) * inv
    return (u >= 0) & (v >= 0) & (u + v < 1)


def _paint(img, mask, color):
    img[mask] = np.asarray(color)  # float colours truncate on assignment, like upstream


def _draw_object(img, x, y, type_, color):
    c = _COLORS[color]
    if type_ == codes.WALL:
        _paint(img, _rect(x, y, 0, 1, 0, 1), c)
    elif type_ == codes.GOAL:
        _paint(img, _rect(x, y, 0, 1, 0, 1), c)
    elif type_ == codes.FLOOR:
        _paint(img, _rect(x, y, 0.031, 1, 0.031, 1), c / 2)
    elif type_ == codes.LAVA:
        _paint(img, _rect(x, y, 0, 1, 0, 1), (255, 128, 0))
        for i in range(3):
            ylo, yhi = 0.3 + 0.2 * i, 0.4 + 0.2 * i
            _paint(img, _line(x, y, 0.1, ylo, 0.3, yhi, 0.03), (0, 0, 0))
            _paint(img, _line(x, y, 0.3, yhi, 0.5, ylo, 0.03), (0, 0, 0))
            _paint(img, _line(x, y, 0.5, ylo, 0.7, yhi, 0.03), (0, 0, 0))
            _paint(img, _line(x, y, 0.7, yhi, 0.9, ylo, 0.03), (0, 0, 0))
    elif type_ == codes.DOOR_OPEN:
        _paint(img, _rect(x, y, 0.88, 1.00, 0.00, 1.00), c)
        _paint(img, _rect(x, y, 0.92, 0.96, 0.04, 0.96), (0, 0, 0))
    elif type_ == codes.DOOR_LOCKED:
        _paint(img, _rect(x, y, 0.00, 1.00, 0.00, 1.00), c)
        _paint(img, _rect(x, y, 0.06, 0.94, 0.06, 0.94), 0.45 * c)
        _paint(img, _rect(x, y, 0.52, 0.75, 0.50, 0.56), c)
    elif type_ == codes.DOOR_CLOSED:
        _paint(img, _rect(x, y, 0.00, 1.00, 0.00, 1.00), c)
        _paint(img, _rect(x, y, 0.04, 0.96, 0.04, 0.96), (0, 0, 0))
        _paint(img, _rect(x, y, 0.08, 0.92, 0.08, 0.92), c)
        _paint(img, _rect(x, y, 0.12, 0.88, 0.12, 0.88), (0, 0, 0))
        _paint(img, _circle(x, y, 0.75, 0.50, 0.08), c)
    elif type_ == codes.KEY:
        _paint(img, _rect(x, y, 0.50, 0.63, 0.31, 0.88), c)
        _paint(img, _rect(x, y, 0.38, 0.50, 0.59, 0.66), c)
        _paint(img, _rect(x, y, 0.38, 0.50, 0.81, 0.88), c)
        _paint(img, _circle(x, y, 0.56, 0.28, 0.190), c)
        _paint(img, _circle(x, y, 0.56, 0.28, 0.064), (0, 0, 0))
    elif type_ == codes.BALL:
        _paint(img, _circle(x, y, 0.5, 0.5, 0.31), c)
    elif type_ == codes.BOX:
        _paint(img, _rect(x, y, 0.12, 0.88, 0.12, 0.88), c)
        _paint(img, _rect(x, y, 0.18, 0.82, 0.18, 0.82), (0, 0, 0))
        _paint(img, _rect(x, y, 0.16, 0.84, 0.47, 0.53), c)
    elif type_ != codes.EMPTY:
        raise ValueError(f"no renderer for type {type_}")


def render_tile(type_, color, agent=False, highlight=False, tile=TILE, agent_dir=3):
    n = tile * SUBDIVS
    x, y = _sample_grid(n)
    img = np.zeros((n, n, 3), dtype=np.uint8)
    _paint(img, _rect(x, y, 0, 0.031, 0, 1), (100, 100, 100))
    _paint(img, _rect(x, y, 0, 1, 0, 0.031), (100, 100, 100))
    _draw_object(img, x, y, type_, color)
    if agent:  # agent_dir is 3 in a POV frame: rotate sample points by -theta about the centre, then test
        theta = 0.5 * math.pi * agent_dir
        xs, ys = x - 0.5, y - 0.5
        x2 = 0.5 + xs * math.cos(-theta) - ys * math.sin(-theta)
        y2 = 0.5 + ys * math.cos(-theta) + xs * math.sin(-theta)
        _paint(img, _triangle(x2, y2, (0.12, 0.19), (0.87, 0.50), (0.12, 0.81)), (255, 0, 0))
    if highlight:
        blend = img + 0.30 * (np.array((255, 255, 255), dtype=np.uint8) - img)
        img[:, :, :] = blend.clip(0, 255).astype(np.uint8)
    small = img.reshape(tile, SUBDIVS, tile, SUBDIVS, 3).mean(axis=3).mean(axis=1)
    out = np.zeros((tile, tile, 3), dtype=np.uint8)
    out[:, :, :] = small
    return out


_cache: dict = {}


def build_atlas(tile=TILE):
    """[128, tile, tile, 3] u8 indexed by packed code / special slot (see module docstring)."""
    if tile in _cache:
        return _cache[tile]
    atlas = np.zeros((N_TILES, tile, tile, 3), dtype=np.uint8)
    atlas[0] = render_tile(codes.EMPTY, 0, highlight=False, tile=tile)         # invisible cell
    atlas[codes.EMPTY] = render_tile(codes.EMPTY, 0, highlight=True, tile=tile)
    atlas[codes.AGENT] = render_tile(codes.EMPTY, 0, agent=True, highlight=True, tile=tile)
    for color in range(6):
        for t in codes.VALID_TYPES:
            if t == codes.EMPTY:
                continue
            fixed = {codes.GOAL: 1, codes.LAVA: 0}.get(t, color)  # Goal() is always green, Lava() always red
            atlas[t | (color << 4)] = render_tile(t, fixed, highlight=True, tile=tile)
        for t in (codes.KEY, codes.BALL, codes.BOX):                               # carried object under the agent
            atlas[(t + 8) | (color << 4)] = render_tile(t, color, agent=True, highlight=True, tile=tile)
    _cache[tile] = atlas
    return atlas
