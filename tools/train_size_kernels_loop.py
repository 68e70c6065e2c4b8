#!/usr/bin/env python
"""Minimal driver for ncu captures of the kernels at TRAINING sizes (config 3: 4096 envs x 128 steps): the policy-input
render (render_f32, 4096 frames, CTA-per-frame mode), the minibatch renders (16 384 gathered frames, f32 and u8), the fused
policy transition on symbolic observations (env_kernel_sym<2>) and GAE (gae_tile_kernel, T = 128, N = 4096)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ppo-2dgrid_b200"))
import torch
from merlin_b200 import BatchedMerlinEnv, gae

dev = "cuda:0"
N, T = 4096, 128
env = BatchedMerlinEnv(N, width=16, height=16, device=dev, generate=("mediumhard", 1, 65536), want_rgb=False)
env.reset(frames=False)
logits = torch.randn(N, 3, device=dev)
value = torch.randn(N, device=dev)
io = env.make_policy_io(logits, value)
rows = []
for t in range(8):
    _, _, _, _, info = env.policy_step(io, frames=False)
    rows.append(info["obs_symbolic"].clone())
sym = torch.cat(rows)                                   # 32 768 stored symbolic observations
pin = torch.empty((N, 14, 14, 48), dtype=torch.float32, device=dev)
mb = torch.empty((16384, 14, 14, 48), dtype=torch.float32, device=dev)
mb8 = torch.empty((16384, 14, 14, 48), dtype=torch.uint8, device=dev)
idx = torch.randperm(sym.shape[0], device=dev)[:16384].contiguous()
for _ in range(3):
    env.render(rows[-1], None, out=pin, blocked=True, dtype=torch.float32, normalise="divide")
    env.render(sym, idx, out=mb, blocked=True, dtype=torch.float32, normalise="divide")
    env.render(sym, idx, out=mb8, blocked=True)
rew, val = torch.rand(T, N, device=dev), torch.randn(T, N, device=dev)
done, last = (torch.rand(T, N, device=dev) < 0.01).float(), torch.randn(N, device=dev)
for _ in range(3):
    gae(rew, val, done, last)
torch.cuda.synchronize()
print("done")
