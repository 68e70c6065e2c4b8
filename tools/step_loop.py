#!/usr/bin/env python
"""Minimal driver for ncu: N envs (argv[1]), K steps (argv[2]), mode rgb|symbolic (argv[3])."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ppo-2dgrid_b200"))
import torch
from merlin_b200 import BatchedMerlinEnv, layouts
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rgb = (sys.argv[3] if len(sys.argv) > 3 else "rgb") == "rgb"
cells, agent = layouts.generate("mediumhard", 16, range(777_000_000, 777_000_000 + 1024))
env = BatchedMerlinEnv(N, cells, agent, width=16, height=16, device="cuda:0", want_rgb=rgb, want_symbolic=not rgb)
env.reset()
acts = torch.randint(0, 3, (16, N), device="cuda:0")
for i in range(K):
    env.step(acts[i % 16])
torch.cuda.synchronize()
print("done", N, K)
