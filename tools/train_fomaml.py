#!/usr/bin/env python
"""FOMAML meta-training on task-batched device rollouts (BASELINE config 4: mediumhard, 1000 iterations x 32 tasks x
k_steps = 256 support + 256 query; with torchrun the tasks of every meta-batch are sharded over the ranks and the
meta-gradient is all-reduced once per iteration).  Mirrors fomaml/fomaml_train.py:100-121 (seed 777, task seeds
`np.random.choice(range(100000), tasks, replace=False)`, lr_inner 0.01, lr_outer 3e-4).

    python tools/train_fomaml.py --iterations 50 --out profiles/r01_fomaml.json
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/train_fomaml.py ...
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ppo-2dgrid_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--difficulty", default="mediumhard")
    ap.add_argument("--iterations", type=int, default=1000)
    ap.add_argument("--tasks-per-batch", type=int, default=32)
    ap.add_argument("--k-steps", type=int, default=256)
    ap.add_argument("--seed", type=int, default=777)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--cpu-baseline", action="store_true")
    ap.add_argument("--save", default=None, help="path for the final meta-policy state_dict (.pth, reference format)")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()

    import numpy as np
    import torch
    import torch.distributed as dist

    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from src.fomaml import FOMAML
    from src.scenario_creator.scenario_creator import ScenarioCreator
    from src.utils.utils import set_seed

    set_seed(a.seed)                       # identical numpy stream on every rank -> identical task batches
    torch.backends.cudnn.benchmark = True
    sc = ScenarioCreator()
    fo = FOMAML(sc, lr_inner=0.01, lr_outer=3e-4, difficulty=a.difficulty, device=dev)
    torch.manual_seed(a.seed + 1000 * rank)  # independent action noise per rank (after the identical weight init)

    def batch():
        return [int(s) for s in np.random.choice(range(100000), size=a.tasks_per_batch, replace=False)]

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(a.warmup):
        fo.meta_train_step(batch(), k_support=a.k_steps, k_query=a.k_steps)
    sync()
    hist = []
    t0 = time.perf_counter()
    nxt = batch()
    for it in range(1, a.iterations + 1):
        cur, nxt = nxt, batch()
        fo.prefetch_tasks(nxt)  # the next meta-batch's layouts are built on the host while the GPU runs this one
        loss, rew, steps, stats = fo.meta_train_step(cur, k_support=a.k_steps, k_query=a.k_steps)
        hist.append({"iter": it, "loss": loss, "rew": rew, "steps": steps, "kl": stats.get("kl", 0.0)})
    sync()
    wall = time.perf_counter() - t0
    if rank == 0:
        env_steps = a.iterations * a.tasks_per_batch * 2 * a.k_steps
        out = {"metric": "FOMAML meta-training env-steps/s (support+query rollouts, inner SGD, meta update)",
               "value": env_steps / wall, "unit": "env-steps/s", "n_gpus": world, "wall_s": wall,
               "iterations": a.iterations, "s_per_iteration": wall / a.iterations,
               "config": {"workload": f"configs[3]: FOMAML {a.difficulty} 16x16, {a.tasks_per_batch} tasks x k={a.k_steps} "
                                      f"support + {a.k_steps} query, tasks sharded over {world} GPU(s)",
                          "tasks_per_gpu": -(-a.tasks_per_batch // world), "seed": a.seed,
                          "dtype": "fp32 policy (PyTorch, stacked per-task weights under vmap), u8 frames"},
               "projected_wall_s_1000_iterations": 1000 * wall / a.iterations,
               "history": hist[:: max(1, len(hist) // 20)] + hist[-1:]}
        if a.cpu_baseline:
            from oracle import ppo_ref
            cb = ppo_ref.cpu_fomaml_task_seconds(a.difficulty, 16, a.k_steps)
            per_iter = cb["seconds_per_task"] * a.tasks_per_batch
            out["cpu_baseline"] = {"value": a.tasks_per_batch * 2 * a.k_steps / per_iter, "unit": "env-steps/s",
                                   "cores": cb["torch_threads"], "kind": "port",
                                   "sample": f"one task of one meta-iteration (k={a.k_steps} support + inner SGD step + "
                                             f"k query + backward) = {cb['seconds_per_task']:.2f} s, reference-style "
                                             "serial loop, literal minigrid restatement, torch CPU; x tasks_per_batch",
                                   "s_per_iteration": per_iter, "projected_wall_s_1000_iterations": 1000 * per_iter}
        if a.save:
            torch.save(fo.meta_policy.state_dict(), a.save)
        print(json.dumps(out), flush=True)
        if a.out:
            with open(a.out, "w") as f:
                json.dump(out, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
