"""Seeding and device resolution behind the reference's names (`set_seed`, `get_device`: src/utils/utils.py:5-46).

The env kernels only exist on CUDA; these helpers decide where the actor-critic lives and seed every host-side
generator the learners draw from (python `random`, numpy's legacy global stream, torch CPU + all CUDA devices).
"""
import random

import numpy as np
import torch

_CPU = torch.device("cpu")


def set_seed(seed):
    for seeder in (random.seed, np.random.seed, torch.manual_seed):
        seeder(seed)
    if not torch.cuda.is_available():
        return
    torch.cuda.manual_seed_all(seed)
    cudnn = torch.backends.cudnn
    cudnn.deterministic, cudnn.benchmark = True, False


def _announce(what):
    print(f"Device set to: {what}")


def get_device(device_str="cpu"):
    """Accepts "cpu", "cuda", "cuda:k" or "auto".

    "auto" picks cuda:0 when a GPU is visible and the CPU otherwise (the reference answers "mps" in that case,
    which Linux does not have).  An unavailable CUDA request or an unknown flag degrades to the CPU with the
    reference's warnings — for the networks only; the batched env refuses to be built without a GPU.
    """
    have_gpu = torch.cuda.is_available()
    wanted = device_str
    if wanted == "auto":
        wanted = "cuda:0" if have_gpu else "cpu"
    if wanted == "cpu":
        _announce("CPU")
        return _CPU
    if not wanted.startswith("cuda"):
        print("[WARNING] Unknown device flag, defaulting to CPU")
        return _CPU
    if not have_gpu:
        print("[WARNING] CUDA requested but not available -> using CPU")
        return _CPU
    chosen = torch.device(wanted)
    _announce(f"{torch.cuda.get_device_name(chosen)} ({wanted})")
    return chosen
