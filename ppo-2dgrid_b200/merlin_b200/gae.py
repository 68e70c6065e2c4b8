"""GAE advantages + returns on the device (replaces PPO.compute_gae, reference src/ppo.py:107-120, and the
loop inlined in FOMAML.compute_loss, src/fomaml.py:116-123)."""
from __future__ import annotations

import torch

from . import _lib


def gae(rewards, values, dones, last_value, gamma=0.99, lam=0.95):
    """rewards/values/dones: f32 CUDA tensors [T] or [T, N] (time-major); last_value: python float, 0-dim or [N].

    Returns (adv, returns) with the input shape.  `returns = values + adv` (PPO convention, src/ppo.py:119);
    FOMAML's normalise-then-add convention (src/fomaml.py:126-127) is applied by its caller.
    """
    if not rewards.is_cuda:
        raise RuntimeError("merlin_b200.gae runs on the GPU only (no CPU fallback): move the rollout to a CUDA device")
    squeeze = rewards.dim() == 1
    r = rewards.detach().to(torch.float32).reshape(rewards.shape[0], -1).contiguous()
    v = values.detach().to(torch.float32).reshape(r.shape).contiguous()
    d = dones.detach().to(torch.float32).reshape(r.shape).contiguous()
    T, N = r.shape
    lv = torch.as_tensor(last_value, dtype=torch.float32, device=r.device).reshape(-1)
    if lv.numel() == 1 and N > 1:
        lv = lv.expand(N)
    lv = lv.contiguous()
    if lv.numel() != N:
        raise ValueError(f"last_value has {lv.numel()} entries for {N} environments")
    adv = torch.empty_like(r)
    ret = torch.empty_like(r)
    with torch.cuda.device(r.device):
        stream = torch.cuda.current_stream().cuda_stream
        _lib.check(_lib.load().merlin_gae(r.data_ptr(), v.data_ptr(), d.data_ptr(), lv.data_ptr(), adv.data_ptr(),
                                          ret.data_ptr(), T, N, float(gamma), float(lam), stream))
    if squeeze:
        return adv.reshape(-1), ret.reshape(-1)
    return adv.reshape(rewards.shape), ret.reshape(rewards.shape)
