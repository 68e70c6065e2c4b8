#!/usr/bin/env python
"""BASELINE config 1 on the CPU with the REAL reference learner: `src.ppo.PPO` from /root/reference (imported over
oracle/shim.py; minigrid/gymnasium underneath are the restatement), mediumhard 16x16, seed 777, the ppo_train.py
defaults (batch 2048, minibatch 256, 10 epochs, lr 3e-4, ent 0.05), `--total-steps` env steps.  Only runs where the
reference checkout is mounted (the build container); writes wall-clock and the learning curve as JSON.

    python tools/reference_cpu_run.py --total-steps 50000 --out profiles/r01_reference_ppo_cpu_50k.json
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--total-steps", type=int, default=50000)
    ap.add_argument("--seed", type=int, default=777)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    from oracle import shim
    shim.install()
    import numpy as np
    import torch
    import src.custom_envs.register  # noqa: F401  (reference)
    from src.ppo import PPO
    from src.scenario_creator.scenario_creator import ScenarioCreator
    from src.utils.utils import set_seed

    set_seed(a.seed)
    env = ScenarioCreator(shim.REFERENCE_ROOT + "/src/config/scenario.yaml").create_env("mediumhard")
    env.reset(seed=a.seed)  # the reference leaves the env RNG to OS entropy; pinned here for repeatability
    agent = PPO(env, lr=3e-4, gamma=0.99, lam=0.95, clip_eps=0.2, update_epochs=10, batch_size=2048, minibatch_size=256,
                vf_coef=0.5, ent_coef=0.05, device="cpu")
    steps, log, t_roll, t_upd = 0, [], 0.0, 0.0
    t0 = time.perf_counter()
    while steps < a.total_steps:
        ta = time.perf_counter()
        lv = agent.collect_rollouts()
        tb = time.perf_counter()
        m = agent.update(lv)
        tc = time.perf_counter()
        t_roll += tb - ta
        t_upd += tc - tb
        steps += agent.batch_size
        rets = agent.episode_returns[-20:]
        log.append({"steps": steps, "mean_return_last20": float(np.mean(rets)) if rets else 0.0, "kl": m["kl"],
                    "entropy": m["entropy"]})
    wall = time.perf_counter() - t0
    out = {"what": "reference src.ppo.PPO (real code, over the minigrid restatement) on the CPU, BASELINE config 1",
           "total_steps": steps, "wall_s": wall, "steps_per_s": steps / wall, "rollout_s": t_roll, "update_s": t_upd,
           "torch_threads": torch.get_num_threads(), "host_cores": len(os.sched_getaffinity(0)),
           "episodes": len(agent.episode_returns), "seed": a.seed, "log": log}
    print(json.dumps({k: v for k, v in out.items() if k != "log"}))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
