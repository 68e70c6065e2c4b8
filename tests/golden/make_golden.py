"""Generate the committed golden fixtures in tests/golden/ by running the REAL reference code
(/root/reference, imported over oracle/shim.py) -- run in the build container only:

    python tests/golden/make_golden.py

What is real reference code here: every `_gen_grid` (src/custom_envs/*.py), `ScenarioCreator.create_env`
(src/scenario_creator/scenario_creator.py:35-57), `ThreeActionWrapper`, `StuckPenaltyWrapper`,
`PPO.compute_gae` (src/ppo.py:107-120) and `compute_gae_standard` (src/utils/utils_rl.py:11-30, the
loop FOMAML.compute_loss inlines at src/fomaml.py:116-123).
What is NOT real: `minigrid`/`gymnasium` underneath them are the restatement in
oracle/minigrid_restated.py (upstream is not installable here) -- so these fixtures pin the
reference-owned logic exactly and the upstream semantics only as restated.

Fixture format (all .npz, deflate-compressed):
  layouts_<difficulty>_<size>.npz  seeds[L], enc u8[L,W,H,3] (Grid.encode()), agent i32[L,3] (x,y,dir)
  trace_<name>.npz                 per-episode layouts + per-step actions/obs/reward/flags/pose
  stuck_trace.npz                  StuckPenaltyWrapper rewards + info["stuck"]
  gae.npz                          inputs + (adv, returns) from the reference GAE loops
  ppo_config1_rollout.npz          one reference PPO rollout (2048 steps, seed 777): layouts, actions, rewards, dones,
                                   values, frame CRCs and the reference's own GAE output
"""
from __future__ import annotations

import os
import sys
import types
import zlib
from collections import deque

import numpy as np

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), "..", ".."))
sys.path.insert(0, ROOT)
OUT = os.path.dirname(os.path.abspath(__file__))

from oracle import shim  # noqa: E402

shim.install()

import torch  # noqa: E402
import src.custom_envs.register  # noqa: E402,F401  (reference, over the shim)
from src.scenario_creator.scenario_creator import ScenarioCreator  # noqa: E402
from src.wrappers.stuck_penalty_wrapper import StuckPenaltyWrapper  # noqa: E402
from src.ppo import PPO  # noqa: E402
from src.utils.utils_rl import compute_gae_standard  # noqa: E402

YAML = os.path.join(shim.REFERENCE_ROOT, "src", "config", "scenario.yaml")
DIFFS = ["easy", "medium", "mediumhard", "hard", "hardest"]


def sized_creator(size):
    """ScenarioCreator whose yaml `size` is overridden (scale-generalisation knob, base_env.py:17)."""
    sc = ScenarioCreator(YAML)
    for cfg in sc.config["difficulties"].values():
        cfg["params"]["size"] = size
    return sc


def snapshot(env):
    u = env.unwrapped
    return u.grid.encode().copy(), np.array([u.agent_pos[0], u.agent_pos[1], u.agent_dir], dtype=np.int32)


def gen_layouts():
    cases = [(d, 16) for d in DIFFS] + [("hard", 8), ("hard", 32), ("mediumhard", 32), ("hardest", 24),
                                        ("mediumhard", 8)]
    for diff, size in cases:
        sc = sized_creator(size)
        seeds = list(range(48)) + [123, 777, 1000, 100000, 200000, 300000, 777000000, 777000001]
        if size != 16:
            seeds = seeds[:16]
        encs, agents = [], []
        env = sc.create_env(diff)
        for s in seeds:
            env.reset(seed=s)
            e, a = snapshot(env)
            encs.append(e)
            agents.append(a)
        np.savez_compressed(
            os.path.join(OUT, f"layouts_{diff}_{size}.npz"),
            seeds=np.array(seeds, dtype=np.int64), enc=np.stack(encs), agent=np.stack(agents),
        )
        print("layouts", diff, size, len(seeds))


def bfs_actions(env):
    """Shortest turn/forward action list from the current pose to the goal (3-action ids)."""
    u = env.unwrapped
    W, H = u.width, u.height
    goal = None
    free = np.zeros((W, H), dtype=bool)
    for i in range(W):
        for j in range(H):
            c = u.grid.get(i, j)
            free[i, j] = c is None or c.type == "goal"
            if c is not None and c.type == "goal":
                goal = (i, j)
    start = (int(u.agent_pos[0]), int(u.agent_pos[1]), int(u.agent_dir))
    vec = [(1, 0), (0, 1), (-1, 0), (0, -1)]
    prev = {start: None}
    q = deque([start])
    end = None
    while q:
        s = q.popleft()
        x, y, d = s
        if (x, y) == goal:
            end = s
            break
        nxt = [((x, y, (d - 1) % 4), 0), ((x, y, (d + 1) % 4), 1)]
        fx, fy = x + vec[d][0], y + vec[d][1]
        if free[fx, fy]:
            nxt.append(((fx, fy, d), 2))
        for ns, a in nxt:
            if ns not in prev:
                prev[ns] = (s, a)
                q.append(ns)
    acts = []
    while prev[end] is not None:
        end, a = prev[end]
        acts.append(a)
    return acts[::-1]


def run_trace(name, diff, size, seed, n_steps, policy, max_steps_override=None, stuck=False):
    sc = sized_creator(size)
    env = sc.create_env(diff)
    if stuck:
        env = StuckPenaltyWrapper(env)
    rng = np.random.default_rng(seed + 17)
    obs, _ = env.reset(seed=seed)
    u = env.unwrapped
    if max_steps_override is not None:
        u.max_steps = max_steps_override
    ep_enc, ep_agent = [], []
    e, a = snapshot(env)
    ep_enc.append(e)
    ep_agent.append(a)
    rec = {k: [] for k in ("action", "obs_rgb", "obs_sym", "reward", "terminated", "truncated", "pose",
                           "episode", "reset_obs_rgb", "reset_obs_sym", "stuck")}
    reset_rgb = [obs.copy()]
    reset_sym = [u.gen_obs()["image"].copy()]
    plan = []
    ep = 0
    for t in range(n_steps):
        if policy == "random":
            act = int(rng.integers(0, 3))
        elif policy == "goal":  # BFS to the goal, with a few random detours
            if not plan:
                plan = bfs_actions(env)
            act = plan.pop(0) if rng.random() > 0.1 else int(rng.integers(0, 2))
            if act in (0, 1) and plan and plan[0] != act:
                plan = []  # detour taken: re-plan next step
        elif policy == "left":
            act = 0
        elif policy == "bump":  # mostly forward (wall banging) with rare turns
            act = 2 if rng.random() > 0.15 else int(rng.integers(0, 2))
        else:
            raise ValueError(policy)
        obs, r, te, tr, info = env.step(act)
        rec["action"].append(act)
        rec["obs_rgb"].append(obs.copy())
        rec["obs_sym"].append(u.gen_obs()["image"].copy())
        rec["reward"].append(float(r))
        rec["terminated"].append(bool(te))
        rec["truncated"].append(bool(tr))
        rec["pose"].append([u.agent_pos[0], u.agent_pos[1], u.agent_dir, u.step_count])
        rec["episode"].append(ep)
        rec["stuck"].append(bool(info.get("stuck", False)))
        if te or tr:
            obs, _ = env.reset()  # like src/ppo.py:96 -- continues the RNG stream
            ep += 1
            plan = []
            e, a = snapshot(env)
            ep_enc.append(e)
            ep_agent.append(a)
            reset_rgb.append(obs.copy())
            reset_sym.append(u.gen_obs()["image"].copy())
    np.savez_compressed(
        os.path.join(OUT, f"trace_{name}.npz"),
        difficulty=diff, size=size, seed=seed, max_steps=u.max_steps, stuck_wrapper=stuck,
        ep_enc=np.stack(ep_enc), ep_agent=np.stack(ep_agent),
        reset_obs_rgb=np.stack(reset_rgb), reset_obs_sym=np.stack(reset_sym),
        action=np.array(rec["action"], dtype=np.int64),
        obs_rgb=np.stack(rec["obs_rgb"]), obs_sym=np.stack(rec["obs_sym"]),
        reward=np.array(rec["reward"], dtype=np.float64),
        terminated=np.array(rec["terminated"]), truncated=np.array(rec["truncated"]),
        pose=np.array(rec["pose"], dtype=np.int32), episode=np.array(rec["episode"], dtype=np.int32),
        stuck=np.array(rec["stuck"]),
    )
    print("trace", name, "steps", n_steps, "episodes", ep + 1, "goals", int(np.sum(rec["terminated"])),
          "trunc", int(np.sum(rec["truncated"])))


def gen_traces():
    for d in DIFFS:
        run_trace(f"{d}_random", d, 16, 123, 160, "random")
        run_trace(f"{d}_goal", d, 16, 777, 160, "goal")
    run_trace("mediumhard_truncate", "mediumhard", 16, 5, 1100, "left")          # 1024 lefts -> truncation
    run_trace("mediumhard_short_horizon", "mediumhard", 16, 9, 200, "random", max_steps_override=37)
    run_trace("hard_32_goal", "hard", 32, 200000, 200, "goal")
    run_trace("mediumhard_8_random", "mediumhard", 8, 3, 120, "random")
    run_trace("stuck_bump", "mediumhard", 16, 42, 200, "bump", stuck=True)
    run_trace("stuck_goal", "medium", 16, 11, 120, "goal", stuck=True)


def gen_gae():
    rng = np.random.default_rng(2024)
    out = {}
    for tag, T, p_done, gamma, lam in [("ppo_2048", 2048, 0.002, 0.99, 0.95), ("ppo_256", 256, 0.05, 0.99, 0.95),
                                       ("fomaml_256", 256, 0.02, 0.995, 0.95), ("alldone_64", 64, 1.0, 0.99, 0.95),
                                       ("t1", 1, 0.0, 0.99, 0.95)]:
        rew = (rng.random(T) < 0.02).astype(np.float32) * rng.random(T).astype(np.float32)
        rew -= (rng.random(T) < 0.1).astype(np.float32) * np.float32(0.1)
        val = rng.normal(0, 0.5, T).astype(np.float32)
        done = (rng.random(T) < p_done).astype(np.float32)
        last = float(np.float32(rng.normal(0, 0.5)))  # like last_val_tensor.item() (src/ppo.py:103)
        me = types.SimpleNamespace(gamma=gamma, lam=lam)
        adv_t, ret_t = PPO.compute_gae(me, torch.tensor(rew), torch.tensor(val), torch.tensor(done), last)
        adv_n, ret_n = compute_gae_standard(rew, val, done, last, gamma=gamma, lam=lam)
        out.update({f"{tag}_rew": rew, f"{tag}_val": val, f"{tag}_done": done, f"{tag}_last": np.float64(last),
                    f"{tag}_gamma": np.float64(gamma), f"{tag}_lam": np.float64(lam),
                    f"{tag}_adv_torch": adv_t.numpy(), f"{tag}_ret_torch": ret_t.numpy(),
                    f"{tag}_adv_numpy": np.asarray(adv_n, dtype=np.float32),
                    f"{tag}_ret_numpy": np.asarray(ret_n, dtype=np.float32)})
        print("gae", tag, "torch-vs-numpy max abs", float(np.max(np.abs(adv_t.numpy() - adv_n))))
    np.savez_compressed(os.path.join(OUT, "gae.npz"), **out)


class _Recorder:
    """Transparent env proxy that snapshots the layout after every reset and the frame CRC of every step."""

    def __init__(self, env):
        self._env = env
        self.enc, self.agent, self.reset_crc, self.step_crc = [], [], [], []

    def __getattr__(self, name):
        return getattr(self._env, name)

    def reset(self, **kw):
        obs, info = self._env.reset(**kw)
        e, a = snapshot(self._env)
        self.enc.append(e)
        self.agent.append(a)
        self.reset_crc.append(zlib.crc32(np.ascontiguousarray(obs).tobytes()))
        return obs, info

    def step(self, action):
        out = self._env.step(action)
        self.step_crc.append(zlib.crc32(np.ascontiguousarray(out[0]).tobytes()))
        return out


def gen_ppo_config1():
    """BASELINE config 1, first iteration: the reference's own PPO.collect_rollouts + compute_gae
    (src/ppo.py:64-120) on mediumhard 16x16 with `set_seed(777)` and the ppo_train.py defaults
    (batch 2048, gamma .99, lambda .95), policy on the CPU.  The env RNG is seeded once (reset(seed=777)) because
    the reference leaves it to OS entropy; every later reset continues that stream, as in training."""
    from src.utils.utils import set_seed
    set_seed(777)
    sc = ScenarioCreator(YAML)
    env = _Recorder(sc.create_env("mediumhard"))
    env.reset(seed=777)
    agent = PPO(env, lr=3e-4, gamma=0.99, lam=0.95, clip_eps=0.2, update_epochs=10, batch_size=2048,
                minibatch_size=256, vf_coef=0.5, ent_coef=0.05, device="cpu")
    n_reset_before = len(env.enc)           # reset(seed) + PPO.__init__'s reset
    last_value = agent.collect_rollouts()
    states, actions, logp, rewards, values, dones = agent.buffer.get()
    adv, ret = agent.compute_gae(rewards, values, dones, last_value)
    np.savez_compressed(
        os.path.join(OUT, "ppo_config1_rollout.npz"),
        ep_enc=np.stack(env.enc), ep_agent=np.stack(env.agent), resets_before_rollout=n_reset_before,
        reset_crc=np.array(env.reset_crc, dtype=np.uint32), step_crc=np.array(env.step_crc, dtype=np.uint32),
        action=actions.numpy().astype(np.int64), reward=rewards.numpy(), done=dones.numpy(),
        value=values.numpy(), logp=logp.numpy(), last_value=np.float64(last_value),
        adv=adv.numpy(), ret=ret.numpy(), gamma=np.float64(0.99), lam=np.float64(0.95),
        first_state_crc=np.uint32(zlib.crc32(states[0].numpy().astype(np.uint8).tobytes())),
        episode_returns=np.array(agent.episode_returns, dtype=np.float64),
        episode_lengths=np.array(agent.episode_lengths, dtype=np.int64),
    )
    print("ppo_config1: episodes", len(agent.episode_returns), "layouts", len(env.enc), "reward sum",
          float(rewards.sum()), "dones", int(dones.sum()))


if __name__ == "__main__":
    if "--only-ppo" in sys.argv:
        gen_ppo_config1()
        sys.exit(0)
    gen_layouts()
    gen_traces()
    gen_gae()
    gen_ppo_config1()
    total = sum(os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT) if f.endswith(".npz"))
    print("total fixture bytes", total)
