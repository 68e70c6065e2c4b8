#!/usr/bin/env python
"""Kernel-level breakdown of one PPO update epoch and one rollout (torch.profiler), config 3 shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ppo-2dgrid_b200")):
    sys.path.insert(0, p)
import torch
from torch.profiler import profile, ProfilerActivity
from src.ppo import PPO
from src.scenario_creator.scenario_creator import ScenarioCreator

torch.backends.cudnn.benchmark = True
N, T, MB = 4096, int(os.environ.get("T", 32)), 16384
env = ScenarioCreator().create_batched_env("mediumhard", N, device="cuda:0", seeds=range(4096))
agent = PPO(env, batch_size=N * T, minibatch_size=MB, update_epochs=1, ent_coef=0.05, use_cuda_graph=False)
lv = agent.collect_rollouts(); agent.update(lv); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    lv = agent.collect_rollouts()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    agent.update(lv)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=60))

# A/B: literal conv1 vs blocked conv1, fwd+bwd on one minibatch and fwd on one rollout batch
s, a, *_ = agent.buffer.get()
xs, acts = s.reshape(-1, 56, 56, 3)[:MB], a.reshape(-1)[:MB]
for blocked in (False, True):
    agent.ac.blocked_first_layer = blocked
    for phase in ("fwd4096", "fwdbwd16384"):
        def run():
            if phase == "fwd4096":
                with torch.no_grad():
                    agent.ac.act(xs[:4096])
            else:
                lp, ent, v = agent.ac.evaluate(xs, acts)
                (lp.mean() + v.mean() + ent.mean()).backward()
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            run()
        e1.record(); torch.cuda.synchronize()
        print(f"blocked={blocked} {phase}: {e0.elapsed_time(e1)/10:.3f} ms")
