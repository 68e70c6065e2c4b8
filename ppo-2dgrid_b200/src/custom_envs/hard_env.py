"""HardEnv: the `hard` MERLIN layout family (reference src/custom_envs/hard_env.py; layout routine restated on
arrays in merlin_b200.layouts, same RNG draw order), stepped by the CUDA env kernels."""
from .base_env import BaseCustomEnv


class HardEnv(BaseCustomEnv):
    difficulty = "hard"
