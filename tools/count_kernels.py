#!/usr/bin/env python
"""Kernels launched per policy step of the learners' rollouts (torch.profiler, eager replay of the rollout body):
FOMAML support (shared weights), FOMAML query (per-task weights under vmap), PPO symbolic-storage rollout.
Prints one JSON line: per path the kernels per step and the kernel names with their counts over `steps` steps.

    python tools/count_kernels.py [--steps 8] [--out profiles/r02_kernels_per_step.json]
"""
from __future__ import annotations

import argparse
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ppo-2dgrid_b200")]


def profile(fn, steps):
    import torch
    from torch.profiler import ProfilerActivity, profile as tprofile
    fn()
    torch.cuda.synchronize()
    with tprofile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    names, micros = collections.Counter(), collections.Counter()
    for ev in prof.events():
        if ev.device_type is not None and "cuda" in str(ev.device_type).lower() and not ev.name.startswith("Memcpy") \
                and not ev.name.startswith("Memset"):
            key = ev.name.split("(")[0][:90]
            names[key] += 1
            micros[key] += float(getattr(ev, "device_time_total", 0.0) or getattr(ev, "cuda_time_total", 0.0) or 0.0)
    total = sum(names.values())
    return {"kernels_total": total, "steps": steps, "kernels_per_step": total / steps,
            "kernel_us_per_step": sum(micros.values()) / steps,
            "by_name": {k: {"launches": c, "us_total": round(micros[k], 1)} for k, c in names.most_common()}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import torch
    from src.fomaml import FOMAML, _stack
    from src.ppo import PPO
    from src.scenario_creator.scenario_creator import ScenarioCreator

    torch.backends.cudnn.benchmark = True
    sc = ScenarioCreator()
    fo = FOMAML(sc, device="cuda:0", difficulty="mediumhard")
    fo.use_cuda_graph = False
    seeds = list(range(32))
    env = fo._task_env(seeds)
    fast = _stack(fo.meta_policy, 32)
    out = {}
    out["fomaml_support"] = profile(lambda: fo.collect_trajectory(env, fo.meta_policy, steps=a.steps), a.steps + 1)
    out["fomaml_query"] = profile(lambda: fo.collect_trajectory(env, fo.meta_policy, steps=a.steps, params=fast), a.steps + 1)
    penv = sc.create_batched_env("mediumhard", 4096, device="cuda:0", layouts="device", seeds=1, n_layouts=4096,
                                 want_symbolic=True)
    agent = PPO(penv, batch_size=4096 * a.steps, minibatch_size=4096, use_cuda_graph=False, obs_storage="symbolic")
    out["ppo_rollout"] = profile(agent.collect_rollouts, a.steps + 1)
    print(json.dumps(out))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
