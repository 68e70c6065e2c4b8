#!/usr/bin/env python
"""BASELINE config 5: scale-generalisation evaluation sweep -- the `hard` scenario at larger grid sizes, 100 unseen tasks
(seeds 200000.., src/sweep_checkpoints.py:90), greedy policy, all tasks of a size in flight at once; with torchrun the
seeds are sharded over the ranks (no collective on the data path; results gathered once at the end).

    python tools/eval_sweep.py --ckpt model.pth --sizes 16,24,32,48,64 --out profiles/r01_eval_sweep.json
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
    ap.add_argument("--ckpt", default=None, help="state_dict .pth in the reference's format; default: random init")
    ap.add_argument("--difficulty", default="hard")
    ap.add_argument("--sizes", default="16,24,32,48,64")
    ap.add_argument("--tasks", type=int, default=100)
    ap.add_argument("--first-seed", type=int, default=200000)
    ap.add_argument("--adapt-steps", type=int, default=0,
                    help="> 0: few-shot evaluation -- every task first takes this many inner SGD steps from the checkpoint "
                         "(FOMAML.few_shot_evaluate, src/distribution_over_tasks.py:132-209) before its greedy episode")
    ap.add_argument("--k-support", type=int, default=256)
    ap.add_argument("--lr-inner", type=float, default=0.01)
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
    from src import parallel
    from src.actor_critic import CNNActorCritic
    from src.evaluation import evaluate_seeds

    torch.manual_seed(777)
    policy = CNNActorCritic((56, 56, 3), 3).to(dev)
    if a.ckpt:
        policy.load_state_dict(torch.load(a.ckpt, map_location=dev))
    parallel.broadcast_parameters(policy)
    seeds = list(range(a.first_seed, a.first_seed + a.tasks))
    mine = parallel.shard(seeds)
    rows = []
    for size in [int(s) for s in a.sizes.split(",")]:
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        if not mine:
            r, n, g = (np.zeros(0),) * 3
        elif a.adapt_steps > 0:
            from src.fomaml import FOMAML
            from src.scenario_creator.scenario_creator import ScenarioCreator
            sc = ScenarioCreator()
            sc.config["difficulties"][a.difficulty]["params"]["size"] = size
            fo = FOMAML(sc, lr_inner=a.lr_inner, device=dev, difficulty=a.difficulty)
            fo.meta_policy.load_state_dict(policy.state_dict())
            r, n, g = fo.few_shot_evaluate(mine, k_support=a.k_support, adapt_steps=a.adapt_steps)
        else:
            r, n, g = evaluate_seeds(policy, a.difficulty, size, mine, device=dev)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        if world > 1:
            parts = [None] * world
            dist.all_gather_object(parts, (r.tolist(), n.tolist(), g.tolist(), dt))
            r = np.concatenate([np.asarray(p[0]) for p in parts])
            n = np.concatenate([np.asarray(p[1]) for p in parts])
            g = np.concatenate([np.asarray(p[2]) for p in parts])
            dt = max(p[3] for p in parts)
        rows.append({"size": size, "max_steps": 4 * size * size, "tasks": len(r), "mean_return": float(np.mean(r)),
                     "mean_steps": float(np.mean(n)), "success_rate": float(np.mean(g)), "seconds": dt,
                     "episode_steps_per_s": float(np.sum(n)) / dt})
        if rank == 0:
            print(json.dumps(rows[-1]), flush=True)
    if rank == 0 and a.out:
        with open(a.out, "w") as f:
            json.dump({"what": f"config 5: {a.difficulty} scale-generalisation sweep, greedy policy, seeds {a.first_seed}..",
                       "n_gpus": world, "checkpoint": a.ckpt or "random init",
                       "mode": f"few-shot: {a.adapt_steps} inner step(s) on {a.k_support} support transitions, lr {a.lr_inner}"
                               if a.adapt_steps else "zero-shot", "rows": rows}, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
