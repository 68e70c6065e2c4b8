#!/usr/bin/env python
"""Wall-clock split of one FOMAML meta-iteration (config 4 shapes): env setup, support / query rollouts, losses, grads."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "ppo-2dgrid_b200")]
import torch
from src.fomaml import FOMAML, _stack
from src.scenario_creator.scenario_creator import ScenarioCreator
torch.backends.cudnn.benchmark = True
fo = FOMAML(ScenarioCreator(), device="cuda:0", difficulty="mediumhard")
seeds = list(range(32))
for _ in range(3): fo.meta_train_step(seeds, 256, 256)
torch.cuda.synchronize()
def t(f):
    torch.cuda.synchronize(); t0=time.perf_counter(); r=f(); torch.cuda.synchronize(); return r, (time.perf_counter()-t0)*1e3
env, t_env = t(lambda: fo._task_env(seeds))
sup, t_sup = t(lambda: fo.collect_trajectory(env, fo.meta_policy, steps=256))
fast = _stack(fo.meta_policy, 32)
names = [n for n,_ in fo.meta_policy.named_parameters()]
(loss,_), t_loss = t(lambda: fo.compute_loss(sup, fo.meta_policy, params=fast))
fast2, t_inner = t(lambda: fo._inner_step(fast, loss, names, 0.01))
q, t_q = t(lambda: fo.collect_trajectory(env, fo.meta_policy, steps=256, params=fast2))
(ql,_), t_ql = t(lambda: fo.compute_loss(q, fo.meta_policy, params=fast2))
g, t_g = t(lambda: torch.autograd.grad(ql.sum(), [fast2[n] for n in names]))
print(f"task_env {t_env:.1f} ms | support rollout {t_sup:.1f} | support loss fwd {t_loss:.1f} | inner grad+step {t_inner:.1f} | query rollout {t_q:.1f} | query loss fwd {t_ql:.1f} | query grad {t_g:.1f}")
