"""merlin_b200 -- host layer of the B200-native MERLIN rollout hot path (see ../../DESIGN.md).

`BatchedMerlinEnv` (fused step + observation kernels), `gae` (GAE/returns kernel), `layouts` (host-side
`_gen_grid` for the five difficulties), `tiles` (RGB tile atlas), `codes` (packed cell codes).
Importing this package does not need a GPU; creating an env or calling `gae` does.
"""
from . import codes, layouts, tiles  # noqa: F401
from .env import BatchedMerlinEnv, OBS_SHAPE, SYM_SHAPE  # noqa: F401
from .gae import gae  # noqa: F401
from ._lib import set_kernel_choice, set_observation_path  # noqa: F401

__all__ = ["BatchedMerlinEnv", "gae", "codes", "layouts", "tiles", "OBS_SHAPE", "SYM_SHAPE"]
