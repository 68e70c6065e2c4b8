#!/usr/bin/env python
"""Minimal driver for ncu: N envs (argv[1]), K steps (argv[2]), mode rgb|symbolic (argv[3]), layout pool size (argv[4],
default 1024), "stagger" (argv[5]) = random episode clocks like bench.py (steady-state auto-resets onto new layouts)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ppo-2dgrid_b200"))
import torch
from merlin_b200 import BatchedMerlinEnv, layouts
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
rgb = (sys.argv[3] if len(sys.argv) > 3 else "rgb") == "rgb"
L = int(sys.argv[4]) if len(sys.argv) > 4 else 1024
if L > 8192:  # the pool of bench.py: generated on the device here (same distribution; this driver only feeds ncu)
    env = BatchedMerlinEnv(N, width=16, height=16, device="cuda:0", want_rgb=rgb, want_symbolic=not rgb,
                           generate=("mediumhard", 777, L))
else:
    cells, agent = layouts.generate("mediumhard", 16, range(777_000_000, 777_000_000 + L))
    env = BatchedMerlinEnv(N, cells, agent, width=16, height=16, device="cuda:0", want_rgb=rgb, want_symbolic=not rgb)
env.reset(frames=rgb)
if len(sys.argv) > 5 and sys.argv[5] == "stagger":
    env.stagger_episode_clocks(seed=777)
acts = torch.randint(0, 3, (16, N), device="cuda:0")
for i in range(K):
    env.step(acts[i % 16], frames=rgb)
torch.cuda.synchronize()
print("done", N, K, env.step_kernel() if rgb else "symbolic")
