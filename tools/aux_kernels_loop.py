#!/usr/bin/env python
"""Minimal driver for ncu captures of the auxiliary kernels: GAE (T=64, N=1M), render (262144 frames from stored symbolic
observations, both layouts, with a row gather) and on-device layout generation (65536 mediumhard layouts)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ppo-2dgrid_b200"))
import torch
from merlin_b200 import BatchedMerlinEnv, gae

dev = "cuda:0"
N = 262144
env = BatchedMerlinEnv(N, width=16, height=16, device=dev, generate=("mediumhard", 1, 65536), want_symbolic=True)
obs, sym = env.reset()
acts = torch.randint(0, 3, (N,), device=dev)
env.step(acts)
sym = env.obs_symbolic.clone()
idx = torch.randperm(N, device=dev)
for _ in range(3):
    env.render(sym, idx)
    env.render(sym, idx, blocked=True)
T, M = 64, 1 << 20
rew, val = torch.rand(T, M, device=dev), torch.randn(T, M, device=dev)
done, last = (torch.rand(T, M, device=dev) < 0.01).float(), torch.randn(M, device=dev)
for _ in range(3):
    gae(rew, val, done, last)
env.generate_layouts("mediumhard", 2, 65536)
torch.cuda.synchronize()
print("done")
# float32 minibatch frames (render_f32_kernel) and the symbolic-only step kernel on the row-parallel gen_obs
out = torch.empty((N, 14, 14, 48), dtype=torch.float32, device=dev)
for _ in range(3):
    env.render(sym, idx, out=out, blocked=True, dtype=torch.float32)
del out
env2 = BatchedMerlinEnv(1 << 20, width=16, height=16, device=dev, generate=("mediumhard", 3, 8192), want_rgb=False)
env2.reset()
a2 = torch.randint(0, 3, (4, 1 << 20), device=dev)
for i in range(6):
    env2.step(a2[i % 4])
torch.cuda.synchronize()
print("done 2")
