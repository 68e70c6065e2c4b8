#!/usr/bin/env python
"""BASELINE configs[1]: batched env step + gen_obs (+RGB render) throughput sweep, mediumhard 16x16, 4k..1M envs on one
B200, random actions, synthetic seeded layouts.  For each N: CUDA-event time over K consecutive step launches after
warm-up (actions pre-generated on the device), env-steps/s, achieved algorithmic GB/s and the fraction of the measured
HBM copy peak.  RGB mode (9 710 B/step) is the headline; symbolic-only mode (449 B/step) is reported beside it.
Small batches fit in the 126 MB L2, so for those an L2 flush (a 256 MB memset) precedes every timed launch and the
launches are timed one by one (`l2: flushed`); large batches stream far more than L2 per launch.

    python tools/sweep.py --out profiles/r01_sweep.json
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ppo-2dgrid_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", default="4096,16384,65536,262144,1048576")
    ap.add_argument("--steps", type=int, default=512)
    ap.add_argument("--warmup", type=int, default=32)
    ap.add_argument("--grid", type=int, default=16)
    ap.add_argument("--difficulty", default="mediumhard")
    ap.add_argument("--modes", default="rgb,symbolic")
    ap.add_argument("--compact", action="store_true")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import torch
    from merlin_b200 import BatchedMerlinEnv, layouts

    dev = torch.device("cuda", 0)
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    S = a.grid
    cells, agent = layouts.generate(a.difficulty, S, range(777_000_000, 777_000_000 + 8192))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    for N in [int(x) for x in a.sizes.split(",")]:
        for mode in a.modes.split(","):
            rgb = mode == "rgb"
            algo = (56 * 56 * 3 if rgb else 147) + S * S + 32 + 8 + 6
            env = BatchedMerlinEnv(N, cells, agent, width=S, height=S, device=dev, want_rgb=rgb, want_symbolic=not rgb)
            env.reset()
            acts = torch.randint(0, 3, (16, N), device=dev)
            for i in range(a.warmup):
                env.step(acts[i % 16])
            torch.cuda.synchronize()
            # back-to-back launches (what a rollout loop sees)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(a.steps):
                env.step(acts[i % 16])
            e1.record()
            torch.cuda.synchronize()
            ms_b2b = e0.elapsed_time(e1) / a.steps
            # cold-L2 launches, timed individually
            k = min(a.steps, 64)
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(k)]
            for i, (s, e) in enumerate(evs):
                flush.zero_()
                s.record()
                env.step(acts[i % 16])
                e.record()
            torch.cuda.synchronize()
            ms_cold = sorted(s.elapsed_time(e) for s, e in evs)[k // 2]
            # 64 steps replayed from one CUDA graph: the kernel without the host's per-call cost (how PPO runs it)
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                for i in range(64):
                    env.step(acts[i % 16])
            graph.replay()
            torch.cuda.synchronize()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(4):
                graph.replay()
            g1.record()
            torch.cuda.synchronize()
            ms_graph = g0.elapsed_time(g1) / 256
            row = {"envs": N, "mode": mode, "algorithmic_bytes_per_step": algo,
                   "ms_per_launch_back_to_back": ms_b2b, "env_steps_per_s": N / ms_b2b * 1e3,
                   "achieved_gbs": algo * N / ms_b2b / 1e6, "frac_of_hbm_peak": algo * N / ms_b2b / 1e6 / peak,
                   "ms_per_launch_l2_flushed_median": ms_cold, "achieved_gbs_l2_flushed": algo * N / ms_cold / 1e6,
                   "ms_per_launch_cuda_graph": ms_graph, "env_steps_per_s_cuda_graph": N / ms_graph * 1e3,
                   "working_set_mb": algo * N / 1e6}
            rows.append(row)
            if a.compact:
                print(f"N={N:8d} {mode:8s} b2b {ms_b2b*1e3:8.1f} us  {row['env_steps_per_s']:.3e}/s  frac {row['frac_of_hbm_peak']:.3f}  "
                      f"cold {ms_cold*1e3:8.1f} us  graph {ms_graph*1e3:8.1f} us ({N / ms_graph * 1e3:.3e}/s)", flush=True)
            else:
                print(json.dumps(row), flush=True)
            env.close()
            del env, acts
    out = {"what": f"{a.difficulty} {S}x{S} step+gen_obs sweep on 1 GPU", "hbm_peak_gbs": peak, "steps": a.steps,
           "gpu": torch.cuda.get_device_name(0), "rows": rows}
    if a.out:
        with open(a.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
