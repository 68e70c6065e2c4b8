"""layer_init and the stateless GAE helper (reference src/utils/utils_rl.py:6-30)."""
import numpy as np
import torch


def layer_init(layer, std=np.sqrt(2), bias_const=0.0):
    torch.nn.init.orthogonal_(layer.weight, std)
    torch.nn.init.constant_(layer.bias, bias_const)
    return layer


def compute_gae_standard(rewards, values, dones, last_value, gamma=0.99, lam=0.95):
    """Same contract as the reference helper (arrays in, (adv, returns) float32 arrays out), computed by the
    GAE kernel: inputs are staged to the current CUDA device and the results copied back."""
    from merlin_b200 import gae

    dev = torch.device("cuda", torch.cuda.current_device())
    r = torch.as_tensor(np.asarray(rewards, dtype=np.float32), device=dev)
    v = torch.as_tensor(np.asarray(values, dtype=np.float32), device=dev)
    d = torch.as_tensor(np.asarray(dones, dtype=np.float32), device=dev)
    adv, ret = gae(r, v, d, float(last_value), gamma, lam)
    return adv.cpu().numpy(), ret.cpu().numpy()
