"""Packed cell codes shared by the host layer and the CUDA kernels (see include/merlin_b200.h).

One byte per grid cell: bits 3..0 type (minigrid OBJECT_TO_IDX, with 11 = closed door, 12 = locked door,
4 = open door), bits 6..4 colour (minigrid COLOR_TO_IDX).  `Grid.encode()`-style arrays
(`[.., W, H, 3]`, index [x][y], channels type/colour/state) convert losslessly both ways.
"""
from __future__ import annotations

import numpy as np

EMPTY, WALL, FLOOR, DOOR_OPEN, KEY, BALL, BOX, GOAL, LAVA, AGENT, DOOR_CLOSED, DOOR_LOCKED = (
    1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12)
COLOR_NAMES = ("red", "green", "blue", "purple", "yellow", "grey")
CODE_EMPTY = EMPTY
CODE_WALL = WALL | (5 << 4)     # grey wall
CODE_GOAL = GOAL | (1 << 4)     # green goal
CODE_LAVA = LAVA | (0 << 4)     # red lava

VALID_TYPES = (EMPTY, WALL, FLOOR, DOOR_OPEN, KEY, BALL, BOX, GOAL, LAVA, DOOR_CLOSED, DOOR_LOCKED)


def pack(type_, color=0, state=0):
    """(type, colour, state) of minigrid -> packed byte (scalar or arrays)."""
    t = np.asarray(type_, dtype=np.int64)
    c = np.asarray(color, dtype=np.int64)
    s = np.asarray(state, dtype=np.int64)
    t = np.where(t == 0, EMPTY, t)
    t = np.where(t == 4, np.where(s == 1, DOOR_CLOSED, np.where(s == 2, DOOR_LOCKED, DOOR_OPEN)), t)
    c = np.where(t == GOAL, 1, np.where(t == LAVA, 0, np.where(t == EMPTY, 0, c)))  # upstream Goal()/Lava() fix their colour
    return ((t & 0xF) | ((c & 7) << 4)).astype(np.uint8)


def pack_encoding(enc):
    """`Grid.encode()` arrays [L, W, H, 3] (or [W, H, 3]) -> packed row-major cells [L, H*W]."""
    enc = np.asarray(enc)
    if enc.ndim == 3:
        enc = enc[None]
    if enc.ndim != 4 or enc.shape[-1] != 3:
        raise ValueError(f"expected [L, W, H, 3] encodings, got {enc.shape}")
    L, W, H, _ = enc.shape
    rm = enc.transpose(0, 2, 1, 3)  # [L, y, x, 3]
    return np.ascontiguousarray(pack(rm[..., 0], rm[..., 1], rm[..., 2]).reshape(L, H * W))


def unpack_to_encoding(cells, width, height):
    """Packed cells [L, >=H*W] -> `Grid.encode()` arrays [L, W, H, 3]."""
    cells = np.asarray(cells, dtype=np.uint8)
    if cells.ndim == 1:
        cells = cells[None]
    L = cells.shape[0]
    c = cells[:, : width * height].reshape(L, height, width).astype(np.int64)
    t = c & 0xF
    col = (c >> 4) & 7
    state = np.where(t == DOOR_CLOSED, 1, np.where(t == DOOR_LOCKED, 2, 0))
    t = np.where(t >= DOOR_CLOSED, 4, t)
    enc = np.stack([t, col, state], axis=-1).astype(np.uint8)
    return np.ascontiguousarray(enc.transpose(0, 2, 1, 3))
