"""ORACLE (test/bench infrastructure, NOT product code) -- CPU restatement of the reference's PPO iteration, used
only to time "the reference CPU run" beside the GPU path (tools/train_ppo.py --cpu-baseline, bench tooling).

What it follows (reference file:line): one iteration of `ppo/ppo_train.py:145-148` = `PPO.collect_rollouts`
(src/ppo.py:64-105: fresh reset, 2048 single-env steps, batch-1 policy inference, per-step buffer writes, reset on
done) + `PPO.update` (src/ppo.py:122-168: Python GAE loop over 0-dim tensors, unbiased-std advantage normalisation,
10 epochs x 8 minibatches of 256, clipped surrogate + 0.5 value loss - ent_coef entropy, grad-clip 0.5, Adam 3e-4,
six `.item()` reads per minibatch).  Env = the literal minigrid restatement with the reference wrapper stack
(oracle/merlin_ref.make_env); network = the same architecture as src/actor_critic.py (two Nature-CNN trunks), torch CPU.
The real `src.ppo.PPO` cannot be shipped to the GPU box (the reference tree does not travel and minigrid/gymnasium are
not installable), so this port is labelled kind="port" wherever its numbers appear; `tests/golden/make_golden.py` ran the
real class over the import shim to produce the config-1 rollout fixture this port is checked against.
"""
from __future__ import annotations

import time

import numpy as np
import torch
import torch.nn as nn

from . import merlin_ref as mr


def _trunk():
    def ortho(m, std=np.sqrt(2)):
        nn.init.orthogonal_(m.weight, std)
        nn.init.constant_(m.bias, 0.0)
        return m
    return nn.Sequential(ortho(nn.Conv2d(3, 32, 8, 4)), nn.ReLU(), ortho(nn.Conv2d(32, 64, 4, 2)), nn.ReLU(),
                         ortho(nn.Conv2d(64, 64, 3, 1)), nn.ReLU(), nn.Flatten()), ortho


class RefActorCritic(nn.Module):
    """Separate actor/critic Nature-CNN trunks + 512-wide heads (src/actor_critic.py:5-64)."""

    def __init__(self, n_actions=3):
        super().__init__()
        self.pi_trunk, ortho = _trunk()
        self.v_trunk, _ = _trunk()
        self.pi_head = nn.Sequential(ortho(nn.Linear(576, 512)), nn.ReLU(), ortho(nn.Linear(512, n_actions), 0.01))
        self.v_head = nn.Sequential(ortho(nn.Linear(576, 512)), nn.ReLU(), ortho(nn.Linear(512, 1), 1.0))

    def dist_value(self, obs_nhwc):
        x = obs_nhwc.permute(0, 3, 1, 2) / 255.0
        return (torch.distributions.Categorical(logits=self.pi_head(self.pi_trunk(x))),
                self.v_head(self.v_trunk(x)).squeeze(-1))


def run_iteration(env, net, opt, batch=2048, minibatch=256, epochs=10, gamma=0.99, lam=0.95, clip=0.2, vf=0.5,
                  ent=0.05, max_seconds=None):
    """One collect + update on the CPU.  Returns dict(steps, rollout_s, update_s, minibatches).  `max_seconds` bounds the
    update (the bench samples a prefix of the 80 minibatch steps and scales)."""
    S = torch.zeros((batch, 56, 56, 3))
    A = torch.zeros(batch, dtype=torch.long)
    LP, R, V, D = (torch.zeros(batch) for _ in range(4))
    t0 = time.perf_counter()
    obs, _ = env.reset()
    for t in range(batch):
        x = torch.tensor(obs, dtype=torch.float32).unsqueeze(0)
        with torch.no_grad():
            dist, v = net.dist_value(x)
            a = dist.sample()
            lp = dist.log_prob(a)
        obs, r, te, tr, _ = env.step(a.item())
        S[t], A[t], LP[t], V[t] = x[0], a[0], lp[0], v[0]
        R[t] = torch.tensor(r, dtype=torch.float32)
        D[t] = torch.tensor(te or tr, dtype=torch.float32)
        if te or tr:
            obs, _ = env.reset()
    with torch.no_grad():
        last = net.dist_value(torch.tensor(obs, dtype=torch.float32).unsqueeze(0))[1].item()
    t1 = time.perf_counter()
    adv, ret = mr.gae_ppo(R, V, D, last, gamma, lam)
    adv = (adv - adv.mean()) / (adv.std() + 1e-8)
    done_mb, total_mb = 0, epochs * ((batch + minibatch - 1) // minibatch)
    for _ in range(epochs):
        perm = torch.randperm(batch)
        for s in range(0, batch, minibatch):
            idx = perm[s:s + minibatch]
            dist, v = net.dist_value(S[idx])
            lp = dist.log_prob(A[idx])
            ratio = torch.exp(lp - LP[idx])
            pi_loss = -torch.min(ratio * adv[idx], torch.clamp(ratio, 1 - clip, 1 + clip) * adv[idx]).mean()
            v_loss = ((v - ret[idx]) ** 2).mean()
            e = dist.entropy().mean()
            loss = pi_loss + vf * v_loss - ent * e
            opt.zero_grad(set_to_none=True)
            loss.backward()
            gn = torch.nn.utils.clip_grad_norm_(net.parameters(), 0.5)
            opt.step()
            _ = (pi_loss.item(), v_loss.item(), e.item(), (LP[idx] - lp).mean().item(),
                 ((ratio - 1).abs() > clip).float().mean().item(), gn.item())
            done_mb += 1
            if max_seconds is not None and time.perf_counter() - t1 > max_seconds:
                break
        else:
            continue
        break
    t2 = time.perf_counter()
    return {"steps": batch, "rollout_s": t1 - t0, "update_s": (t2 - t1) * total_mb / max(done_mb, 1),
            "update_measured_s": t2 - t1, "minibatches_run": done_mb, "minibatches_total": total_mb}


def cpu_ppo_sps(difficulty="mediumhard", size=16, seed=777, update_budget_s=20.0, threads=None):
    """Steps/s of one reference-style PPO iteration on this host (all torch CPU threads unless `threads`)."""
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(seed)
    env = mr.make_env(difficulty, size=size)
    env.reset(seed=seed)
    net = RefActorCritic(3)
    opt = torch.optim.Adam(net.parameters(), lr=3e-4)
    res = run_iteration(env, net, opt, max_seconds=update_budget_s)
    res["steps_per_s"] = res["steps"] / (res["rollout_s"] + res["update_s"])
    res["torch_threads"] = torch.get_num_threads()
    return res


# ---- FOMAML: one task of a meta-iteration, reference style -------------------------------------------------------
def _fomaml_traj(env, net, k, seed):
    """src/fomaml.py:54-108 -- k single-env steps, batch-1 inference, `reset(seed=seed)` after every done."""
    obs, _ = env.reset(seed=seed)
    O, A, LP, V, R, D = [], [], [], [], [], []
    for _ in range(k):
        x = torch.tensor(obs, dtype=torch.float32).unsqueeze(0)
        with torch.no_grad():
            dist, v = net.dist_value(x)
            a = dist.sample()
        obs, r, te, tr, _ = env.step(a.item())
        O.append(x); A.append(a); LP.append(dist.log_prob(a)); V.append(v); R.append(r); D.append(te or tr)
        if te or tr:
            obs, _ = env.reset(seed=seed)
    with torch.no_grad():
        last = net.dist_value(torch.tensor(obs, dtype=torch.float32).unsqueeze(0))[1]
    return (torch.cat(O), torch.cat(A), torch.cat(LP), torch.cat(V), torch.tensor(R, dtype=torch.float32),
            torch.tensor(D, dtype=torch.float32), last)


def _fomaml_loss(net, traj, gamma=0.995, lam=0.95, clip=0.2, vf=0.5, ent=0.05):
    """src/fomaml.py:110-156 -- numpy GAE loop, normalise, ret = val + adv_norm, PPO-clip loss."""
    O, A, LP, V, R, D, last = traj
    _, adv_n, ret = mr.gae_fomaml(R.numpy(), V.numpy(), D.numpy(), last.item(), gamma, lam)
    adv_n, ret = torch.tensor(adv_n), torch.tensor(ret)
    dist, v = net.dist_value(O)
    lp = dist.log_prob(A)
    ratio = torch.exp(lp - LP)
    pi = -torch.min(ratio * adv_n, torch.clamp(ratio, 1 - clip, 1 + clip) * adv_n).mean()
    return pi + vf * ((v - ret) ** 2).mean() - ent * dist.entropy().mean()


def cpu_fomaml_task_seconds(difficulty="mediumhard", size=16, k=256, seed=123, threads=None):
    """Wall seconds the reference-style loop spends on ONE task of a meta-iteration (support rollout + inner SGD step +
    query rollout + query backward, src/fomaml.py:167-202) on this host; a meta-iteration is tasks_per_batch of these."""
    if threads:
        torch.set_num_threads(threads)
    torch.manual_seed(seed)
    env = mr.make_env(difficulty, size=size)
    net = RefActorCritic(3)
    inner = torch.optim.SGD(net.parameters(), lr=0.01)
    _fomaml_traj(env, net, 8, seed)  # warm the tile cache / allocator
    t0 = time.perf_counter()
    loss = _fomaml_loss(net, _fomaml_traj(env, net, k, seed))
    inner.zero_grad()
    loss.backward()
    torch.nn.utils.clip_grad_norm_(net.parameters(), 0.5)
    inner.step()
    env.reset(seed=seed)
    q = _fomaml_loss(net, _fomaml_traj(env, net, k, seed))
    net.zero_grad()
    q.backward()
    return {"seconds_per_task": time.perf_counter() - t0, "k": k, "torch_threads": torch.get_num_threads()}
