#!/usr/bin/env python
"""BASELINE config 5: scale-generalisation evaluation sweep -- the `hard` scenario at larger grid sizes, 100 unseen tasks
(seeds 200000.., src/sweep_checkpoints.py:90), greedy policy, every task of a job in flight at once.

Arms (src/sweep_checkpoints.py:58-90, src/distribution_over_tasks.py:71-96,132-209):
    ppo_zero_shot      the PPO checkpoint, greedy                                (--ckpt / --ppo-ckpt)
    fomaml_zero_shot   the FOMAML meta-weights, greedy                           (--fomaml-ckpt)
    fomaml_few_shot    the meta-weights after `--adapt-steps` inner SGD steps per task on `--k-support` transitions
A job = (arm, size) with all `--tasks` seeds.  100 tasks are far too few to fill one GPU, let alone eight, and a job is
4*size^2 sequential steps, so with torchrun the JOBS -- not the tasks -- are spread over the ranks (longest first onto
the least loaded rank); no collective on the data path, results gathered once at the end.

    python tools/eval_sweep.py --ckpt ppo.pth --fomaml-ckpt fomaml.pth --sizes 16,24,32,48,64 --out profiles/r02_eval_sweep.json
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/eval_sweep.py ...
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


def assign_jobs(jobs, world):
    """Longest-processing-time-first: jobs [(cost, ...)] -> list of job lists per rank."""
    loads, plan = [0.0] * world, [[] for _ in range(world)]
    for job in sorted(jobs, key=lambda j: -j[0]):
        r = loads.index(min(loads))
        plan[r].append(job)
        loads[r] += job[0]
    return plan


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ckpt", "--ppo-ckpt", dest="ckpt", default=None,
                    help="PPO state_dict .pth in the reference's format; default: random init")
    ap.add_argument("--fomaml-ckpt", default=None, help="FOMAML meta-policy state_dict: adds the two FOMAML arms")
    ap.add_argument("--difficulty", default="hard")
    ap.add_argument("--sizes", default="16,24,32,48,64")
    ap.add_argument("--tasks", type=int, default=100)
    ap.add_argument("--first-seed", type=int, default=200000)
    ap.add_argument("--adapt-steps", type=int, default=0,
                    help="inner SGD steps of the few-shot arm (FOMAML.few_shot_evaluate); with --fomaml-ckpt default 1; "
                         "without it, > 0 turns the single --ckpt arm into a few-shot arm")
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
    from src.actor_critic import CNNActorCritic
    from src.evaluation import evaluate_seeds

    torch.manual_seed(777)
    torch.backends.cudnn.benchmark = True
    seeds = list(range(a.first_seed, a.first_seed + a.tasks))
    sizes = [int(s) for s in a.sizes.split(",")]

    def load(path):
        policy = CNNActorCritic((56, 56, 3), 3).to(dev)
        if path:
            policy.load_state_dict(torch.load(path, map_location=dev))
        return policy

    arms = []  # (name, checkpoint path, adapt steps)
    if a.fomaml_ckpt:
        arms.append(("ppo_zero_shot", a.ckpt, 0))
        arms.append(("fomaml_zero_shot", a.fomaml_ckpt, 0))
        arms.append(("fomaml_few_shot", a.fomaml_ckpt, max(1, a.adapt_steps)))
    else:
        arms.append(("few_shot" if a.adapt_steps > 0 else "zero_shot", a.ckpt, a.adapt_steps))
    policies = {path: load(path) for path in {arm[1] for arm in arms}}

    # cost model: an episode is at most 4*size^2 sequential policy steps; a few-shot arm adds its support rollouts
    jobs = [(4 * size * size + adapt * (a.k_support + 64), name, path, adapt, size)
            for name, path, adapt in arms for size in sizes]
    mine = assign_jobs(jobs, world)[rank]

    def run(job):
        _, name, path, adapt, size = job
        policy = policies[path]
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        if adapt > 0:
            from src.fomaml import FOMAML
            from src.scenario_creator.scenario_creator import ScenarioCreator
            sc = ScenarioCreator()
            sc.config["difficulties"][a.difficulty].setdefault("params", {})["size"] = size
            fo = FOMAML(sc, lr_inner=a.lr_inner, device=dev, difficulty=a.difficulty, sync_init=False)
            fo.meta_policy.load_state_dict(policy.state_dict())
            r, n, g = fo.few_shot_evaluate(seeds, k_support=a.k_support, adapt_steps=adapt)
        else:
            r, n, g = evaluate_seeds(policy, a.difficulty, size, seeds, device=dev)
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        return {"arm": name, "size": size, "max_steps": 4 * size * size, "tasks": len(r), "mean_return": float(np.mean(r)),
                "mean_steps": float(np.mean(n)), "success_rate": float(np.mean(g)), "seconds": dt,
                "episode_steps_per_s": float(np.sum(n)) / dt, "rank": rank,
                "mode": f"few-shot: {adapt} inner step(s) on {a.k_support} support transitions, lr {a.lr_inner}" if adapt
                        else "zero-shot"}

    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    t_all = time.perf_counter()
    rows = [run(job) for job in mine]
    torch.cuda.synchronize(dev)
    busy = time.perf_counter() - t_all
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, (rows, busy))
        rows = [row for part in parts for row in part[0]]
        busy_all = [part[1] for part in parts]
    else:
        busy_all = [busy]
    wall = time.perf_counter() - t_all
    order = {name: i for i, (name, _, _) in enumerate(arms)}
    rows.sort(key=lambda row: (order[row["arm"]], row["size"]))
    if rank == 0:
        for row in rows:
            print(json.dumps(row), flush=True)
        if a.out:
            with open(a.out, "w") as f:
                json.dump({"what": f"config 5: {a.difficulty} scale-generalisation sweep, greedy policy, seeds {a.first_seed}.., "
                                   f"{a.tasks} tasks per job",
                           "n_gpus": world, "sharding": "(arm, size) jobs over ranks, longest first onto the least loaded rank",
                           "checkpoints": {name: (path or "random init") for name, path, _ in arms},
                           "wall_s_whole_sweep": wall, "busy_s_per_rank": busy_all,
                           "total_episode_steps": float(sum(row["mean_steps"] * row["tasks"] for row in rows)),
                           "rows": rows}, f, indent=1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
