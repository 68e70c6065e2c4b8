#!/usr/bin/env python
"""What a small batch can reach at best: time per launch of a plain device fill of the same number of bytes the step
kernel writes for N envs (9408 B frames), replayed back to back from one CUDA graph like tools/sweep.py replays the step
kernel.  The difference to the step kernel's own time is what its dependent-load chain and frame assembly cost.

    python tools/fill_floor.py [--sizes 4096,8192,16384,24576]
"""
import argparse, json, torch
ap = argparse.ArgumentParser()
ap.add_argument("--sizes", default="4096,8192,16384,24576")
a = ap.parse_args()
dev = torch.device("cuda", 0)
rows = []
for n in [int(s) for s in a.sizes.split(",")]:
    buf = torch.empty(n * 9408, dtype=torch.uint8, device=dev)
    buf.zero_()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(64):
            buf.zero_()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(8):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 512 * 1e3
    rows.append({"envs": n, "bytes": n * 9408, "fill_us_per_launch_cuda_graph": us, "gbs": n * 9408 / us / 1e3})
    print(f"N={n:6d}  {n*9408/1e6:7.1f} MB  fill {us:6.2f} us/launch  {n*9408/us/1e3:7.0f} GB/s", flush=True)
print(json.dumps(rows))
