"""set_seed / get_device with the reference's signatures (src/utils/utils.py:5-46)."""
import random

import numpy as np
import torch


def set_seed(seed):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
        torch.backends.cudnn.deterministic = True
        torch.backends.cudnn.benchmark = False


def get_device(device_str="cpu"):
    """ "cpu" | "cuda" | "cuda:k" | "auto".  The env kernels need CUDA; "auto" without a GPU falls back to the
    CPU torch device for the networks only (the reference returns "mps" there, which does not exist on Linux)."""
    if device_str == "auto":
        device_str = "cuda:0" if torch.cuda.is_available() else "cpu"
    if device_str.startswith("cuda"):
        if torch.cuda.is_available():
            dev = torch.device(device_str)
            print(f"Device set to: {torch.cuda.get_device_name(dev)} ({device_str})")
            return dev
        print("[WARNING] CUDA requested but not available -> using CPU")
        return torch.device("cpu")
    if device_str != "cpu":
        print("[WARNING] Unknown device flag, defaulting to CPU")
    else:
        print("Device set to: CPU")
    return torch.device("cpu")
