"""Shared test helpers: fixture loading and trace replay against any vector-env that follows the
batched semantics (oracle.fast.OracleVecEnv on the CPU, merlin_b200.BatchedMerlinEnv on the GPU)."""
from __future__ import annotations

import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def trace_names():
    return sorted(os.path.basename(p)[len("trace_"):-len(".npz")] for p in glob.glob(os.path.join(GOLDEN, "trace_*.npz")))


def layout_names():
    return sorted(os.path.basename(p)[:-len(".npz")] for p in glob.glob(os.path.join(GOLDEN, "layouts_*.npz")))


def to_np(x):
    if hasattr(x, "detach"):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def replay_trace_autoreset(make_env, tr):
    """N=1, auto-reset on: pool = the episode layouts of the fixture in order (cursor stride 1).
    At done steps the env must return the NEXT episode's reset observation."""
    stuck = bool(tr["stuck_wrapper"])
    env = make_env(num_envs=1, enc=tr["ep_enc"], agent=tr["ep_agent"], max_steps=int(tr["max_steps"]),
                   auto_reset=True, reset_mode="next", stuck_penalty=stuck)
    rgb, sym = env.reset()
    assert np.array_equal(to_np(rgb)[0], tr["reset_obs_rgb"][0])
    assert np.array_equal(to_np(sym)[0], tr["reset_obs_sym"][0])
    ep_ret = 0.0
    for t, a in enumerate(tr["action"]):
        rgb, rew, term, trunc, info = env.step(np.array([a], dtype=np.int64))
        rgb, rew, term, trunc = to_np(rgb), to_np(rew), to_np(term), to_np(trunc)
        sym = to_np(info["obs_symbolic"])
        assert rew.dtype == np.float32
        assert rew[0] == np.float32(tr["reward"][t]), (t, rew[0], tr["reward"][t])
        assert bool(term[0]) == bool(tr["terminated"][t]), t
        assert bool(trunc[0]) == bool(tr["truncated"][t]), t
        ep_ret += float(np.float32(tr["reward"][t]))
        if stuck:
            assert bool(to_np(info["stuck"])[0]) == bool(tr["stuck"][t]), t
        if tr["terminated"][t] or tr["truncated"][t]:
            ep = int(tr["episode"][t]) + 1
            assert np.array_equal(rgb[0], tr["reset_obs_rgb"][ep]), t
            assert np.array_equal(sym[0], tr["reset_obs_sym"][ep]), t
            assert int(to_np(info["episode_length"])[0]) == int(tr["pose"][t][3])
            assert abs(float(to_np(info["episode_return"])[0]) - ep_ret) < 1e-5
            ep_ret = 0.0
        else:
            assert np.array_equal(rgb[0], tr["obs_rgb"][t]), t
            assert np.array_equal(sym[0], tr["obs_sym"][t]), t
            assert int(to_np(info["episode_length"])[0]) == 0
    return env


def replay_trace_manual_reset(make_env, tr):
    """N=1, auto-reset off: step returns the terminal observation, then reset() like src/ppo.py:93-96."""
    stuck = bool(tr["stuck_wrapper"])
    env = make_env(num_envs=1, enc=tr["ep_enc"], agent=tr["ep_agent"], max_steps=int(tr["max_steps"]),
                   auto_reset=False, reset_mode="next", stuck_penalty=stuck)
    rgb, sym = env.reset()
    assert np.array_equal(to_np(rgb)[0], tr["reset_obs_rgb"][0])
    for t, a in enumerate(tr["action"]):
        rgb, rew, term, trunc, info = env.step(np.array([a], dtype=np.int64))
        assert np.array_equal(to_np(rgb)[0], tr["obs_rgb"][t]), t
        assert np.array_equal(to_np(info["obs_symbolic"])[0], tr["obs_sym"][t]), t
        assert to_np(rew)[0] == np.float32(tr["reward"][t])
        assert bool(to_np(term)[0]) == bool(tr["terminated"][t])
        assert bool(to_np(trunc)[0]) == bool(tr["truncated"][t])
        pose = get_pose(env)
        assert tuple(pose[0]) == tuple(tr["pose"][t]), (t, pose[0], tr["pose"][t])
        if tr["terminated"][t] or tr["truncated"][t]:
            rgb, sym = env.reset()
            ep = int(tr["episode"][t]) + 1
            assert np.array_equal(to_np(rgb)[0], tr["reset_obs_rgb"][ep]), t
            assert np.array_equal(to_np(sym)[0], tr["reset_obs_sym"][ep]), t
    return env


def get_pose(env):
    """[N,4] (x, y, dir, step_count) from either kind of vector env."""
    if hasattr(env, "pose_numpy"):
        return env.pose_numpy()
    return np.stack([env.ax, env.ay, env.adir, env.stepc], axis=1)
