"""Drop-in mirror of the reference's `src` package for the rollout hot path (reference src/__init__.py:1-4).

Same importable names and call signatures as borangundogan/PPO-2DGrid; the environment step, observation,
reward shaping and GAE run in libmerlin_b200.so on the GPU, the actor-critic stays PyTorch.
"""
from .actor_critic import MLPActorCritic, CNNActorCritic  # noqa: F401
from .rollout_buffer import RolloutBuffer  # noqa: F401
from .utils.utils import get_device  # noqa: F401
from .utils.utils_rl import layer_init  # noqa: F401
