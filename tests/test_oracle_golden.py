"""CPU tests: pin the oracle (Python literal restatement + C restatement) to the committed fixtures
that were produced by the REAL reference code (tests/golden/make_golden.py), and to the
hand-derivable known answers of SURVEY.md section 8c.6."""
from __future__ import annotations

import numpy as np
import pytest
import torch

import helpers
from oracle import fast, merlin_ref as mr, minigrid_restated as mg


# ---- layouts: restated _gen_grid vs the reference's own _gen_grid ------------------------------
@pytest.mark.parametrize("name", helpers.layout_names())
def test_layouts_match_reference(name):
    fx = helpers.load(name + ".npz")
    _, diff, size = name.split("_")
    env = mr.MerlinEnv(difficulty=diff, size=int(size))
    for k, seed in enumerate(fx["seeds"]):
        env.reset(seed=int(seed))
        assert np.array_equal(env.grid.encode(), fx["enc"][k]), (name, seed)
        assert (env.agent_pos[0], env.agent_pos[1], env.agent_dir) == tuple(fx["agent"][k]), (name, seed)


# ---- traces: restated wrapper stack, step by step ------------------------------------------------
@pytest.mark.parametrize("name", helpers.trace_names())
def test_python_oracle_replays_trace(name):
    tr = helpers.load(f"trace_{name}.npz")
    env = mr.make_env(str(tr["difficulty"]), size=int(tr["size"]), stuck_penalty=bool(tr["stuck_wrapper"]))
    obs, _ = env.reset(seed=int(tr["seed"]))
    env.unwrapped.max_steps = int(tr["max_steps"])
    assert np.array_equal(obs, tr["reset_obs_rgb"][0])
    for t, a in enumerate(tr["action"]):
        obs, r, te, trn, info = env.step(int(a))
        assert np.array_equal(obs, tr["obs_rgb"][t]), t
        assert float(r) == float(tr["reward"][t])
        assert (te, trn) == (bool(tr["terminated"][t]), bool(tr["truncated"][t]))
        if te or trn:
            obs, _ = env.reset()
            assert np.array_equal(obs, tr["reset_obs_rgb"][int(tr["episode"][t]) + 1])


@pytest.mark.parametrize("name", helpers.trace_names())
def test_c_oracle_replays_trace_autoreset(name):
    helpers.replay_trace_autoreset(fast.OracleVecEnv, helpers.load(f"trace_{name}.npz"))


@pytest.mark.parametrize("name", helpers.trace_names())
def test_c_oracle_replays_trace_manual_reset(name):
    helpers.replay_trace_manual_reset(fast.OracleVecEnv, helpers.load(f"trace_{name}.npz"))


# ---- GAE -----------------------------------------------------------------------------------------
GAE_TAGS = ["ppo_2048", "ppo_256", "fomaml_256", "alldone_64", "t1"]


@pytest.mark.parametrize("tag", GAE_TAGS)
def test_gae_restatements_match_reference(tag):
    fx = helpers.load("gae.npz")
    rew, val, done = fx[f"{tag}_rew"], fx[f"{tag}_val"], fx[f"{tag}_done"]
    last, gamma, lam = float(fx[f"{tag}_last"]), float(fx[f"{tag}_gamma"]), float(fx[f"{tag}_lam"])
    adv, ret = mr.gae_ppo(torch.tensor(rew), torch.tensor(val), torch.tensor(done), last, gamma, lam)
    assert np.array_equal(adv.numpy(), fx[f"{tag}_adv_torch"])
    assert np.array_equal(ret.numpy(), fx[f"{tag}_ret_torch"])
    adv_n, _, _ = mr.gae_fomaml(rew, val, done, last, gamma, lam)
    assert np.array_equal(adv_n, fx[f"{tag}_adv_numpy"])
    # C restatement, as a [T,1] rollout: bit-exact against both reference loops
    adv_c, ret_c = fast.gae(rew[:, None], val[:, None], done[:, None], np.float32(last), gamma, lam)
    assert np.array_equal(adv_c[:, 0], fx[f"{tag}_adv_torch"])
    assert np.array_equal(ret_c[:, 0], fx[f"{tag}_ret_torch"])


def test_gae_known_answers():
    # T=1, not done: adv = r + gamma*last - v ; all-done: adv_t = r_t - v_t   
    adv, ret = fast.gae(np.array([[0.5]]), np.array([[0.25]]), np.array([[0.0]]), np.float32(2.0), 0.99, 0.95)
    assert adv[0, 0] == np.float32(np.float32(0.5) + np.float32(0.99 * 2.0) - np.float32(0.25))
    assert ret[0, 0] == np.float32(0.25) + adv[0, 0]
    r = np.linspace(-1, 1, 8, dtype=np.float32)[:, None]
    v = np.linspace(0.5, -0.5, 8, dtype=np.float32)[:, None]
    adv, _ = fast.gae(r, v, np.ones_like(r), np.float32(3.0), 0.99, 0.95)
    assert np.array_equal(adv, r - v)


# ---- known answers for the env -------------------------------------------
def _room(size=16, agent=(1, 1, 0), goal=None, walls=()):
    enc = np.zeros((1, size, size, 3), np.uint8)
    enc[..., 0] = 1
    enc[0, 0, :, :] = enc[0, -1, :, :] = enc[0, :, 0, :] = enc[0, :, -1, :] = (2, 5, 0)
    for (x, y) in walls:
        enc[0, x, y] = (2, 5, 0)
    if goal is not None:
        enc[0, goal[0], goal[1]] = (8, 1, 0)
    return enc, np.array([agent], np.int32)


def test_tile_values():
    at = fast.TileAtlas(8)
    at.ensure([(2, 5, 0), (8, 1, 0)])
    unseen = at.tiles[fast.tile_slot(1, 0, 0, 0, 0)]
    assert unseen[0, 0].tolist() == [55] * 3 and unseen[0, 3].tolist() == [33] * 3
    assert unseen[3, 0].tolist() == [33] * 3 and unseen[3, 3].tolist() == [0] * 3
    lit = at.tiles[fast.tile_slot(1, 0, 0, 0, 1)]
    assert lit[0, 0].tolist() == [114] * 3 and lit[0, 5].tolist() == [99] * 3 and lit[4, 4].tolist() == [76] * 3
    assert np.all(at.tiles[fast.tile_slot(2, 5, 0, 0, 1)] == 146)
    assert np.all(at.tiles[fast.tile_slot(8, 1, 0, 0, 1)] == np.array([76, 255, 76], np.uint8))
    ag = at.tiles[fast.tile_slot(1, 0, 0, 1, 1)]
    assert ag[1:7, 1:8, 0].tolist() == [[76, 76, 115, 115, 76, 76, 76], [76, 76, 175, 175, 76, 76, 76],
                                        [76, 76, 255, 255, 76, 76, 76], [76, 155, 255, 255, 155, 76, 76],
                                        [76, 235, 255, 255, 235, 76, 76], [115, 255, 255, 255, 255, 115, 76]]
    assert np.all(ag[1:, 1:, 1] == 76) and np.all(ag[1:, 1:, 2] == 76)


def test_reward_truncation_and_wall_bump():
    # goal directly ahead: reached on step 1 -> f32(1 - 0.9/1024)
    enc, agent = _room(agent=(1, 1, 0), goal=(2, 1))
    env = fast.OracleVecEnv(1, enc, agent, auto_reset=False)
    env.reset()
    _, r, te, tr, _ = env.step(np.array([2]))
    assert te[0] and not tr[0] and r[0] == np.float32(1 - 0.9 * (1 / 1024))
    # 1024 lefts: truncated exactly at step 1024, reward 0, same heading (1024 % 4 == 0)
    enc, agent = _room(agent=(5, 5, 1))
    env = fast.OracleVecEnv(1, enc, agent, auto_reset=False)
    env.reset()
    for k in range(1024):
        _, r, te, tr, _ = env.step(np.array([0]))
        assert r[0] == 0 and not te[0]
        assert bool(tr[0]) == (k == 1023)
    assert (env.ax[0], env.ay[0], env.adir[0], env.stepc[0]) == (5, 5, 1, 1024)
    # reaching the goal exactly at max_steps: terminated AND truncated, reward f32(0.1)
    enc, agent = _room(agent=(1, 1, 0), goal=(2, 1))
    env = fast.OracleVecEnv(1, enc, agent, auto_reset=False, max_steps=4)
    env.reset()
    for _ in range(3):
        env.step(np.array([0]))
    env.step(np.array([0]))  # 4 lefts (truncated already at step 4, env keeps stepping like upstream)
    enc, agent = _room(agent=(1, 1, 0), goal=(2, 1))
    env = fast.OracleVecEnv(1, enc, agent, auto_reset=False, max_steps=1)
    env.reset()
    _, r, te, tr, _ = env.step(np.array([2]))
    assert te[0] and tr[0] and r[0] == np.float32(1 - 0.9 * (1 / 1))
    # wall bump: pose unchanged, step_count increments
    enc, agent = _room(agent=(1, 1, 2))
    env = fast.OracleVecEnv(1, enc, agent, auto_reset=False)
    env.reset()
    _, r, te, tr, _ = env.step(np.array([2]))
    assert (env.ax[0], env.ay[0], env.adir[0], env.stepc[0]) == (1, 1, 2, 1) and r[0] == 0


def test_view_geometry_and_occlusion():
    # empty room, agent (1,1) facing right: view columns 0,1 out of bounds (walls), column 2 = border row y=0
    enc, agent = _room(agent=(1, 1, 0))
    env = fast.OracleVecEnv(1, enc, agent, auto_reset=False)
    _, sym = env.reset()
    s = sym[0]
    assert s[3, 6].tolist() == [1, 0, 0]          # agent cell shows empty
    assert s[2, 6].tolist() == [2, 5, 0]          # world (1,0): border wall, visible
    assert s[3, 5].tolist() == [1, 0, 0]          # world (2,1)
    assert s[3, 0].tolist() == [1, 0, 0]          # world (7,1)
    assert s[0, 6].tolist() == [0, 0, 0]          # behind the border wall: unseen
    # a full wall across view row 5 hides row 4; a single wall at (3,5) does not (row 4 is lit sideways)
    enc, agent = _room(agent=(5, 8, 0), walls=[(6, y) for y in range(5, 12)])
    env = fast.OracleVecEnv(1, enc, agent, auto_reset=False)
    _, sym = env.reset()
    assert sym[0][3, 5].tolist() == [2, 5, 0] and sym[0][3, 4].tolist() == [0, 0, 0]
    enc, agent = _room(agent=(5, 8, 0), walls=[(6, 8)])
    env = fast.OracleVecEnv(1, enc, agent, auto_reset=False)
    _, sym = env.reset()
    assert sym[0][3, 5].tolist() == [2, 5, 0] and sym[0][3, 4].tolist() == [1, 0, 0]


def test_stuck_penalty_known_answers():
    enc, agent = _room(agent=(1, 1, 2))  # facing the left border wall
    env = fast.OracleVecEnv(1, enc, agent, auto_reset=False, stuck_penalty=True)
    env.reset()
    rewards = [float(env.step(np.array([2]))[1][0]) for _ in range(5)]
    assert rewards == [0.0, 0.0, float(np.float32(-0.1)), float(np.float32(-0.1)), float(np.float32(-0.1))]
    env.step(np.array([0]))  # turn (still stuck: position unchanged)
    env.step(np.array([0]))  # now facing right
    _, r, _, _, info = env.step(np.array([2]))  # moves: counter resets, no penalty
    assert r[0] == 0 and not info["stuck"][0]


def test_exploration_bonus_first_visit_only():
    enc, agent = _room(agent=(1, 1, 0))
    env = fast.OracleVecEnv(1, enc, agent, auto_reset=False, exploration_bonus=0.01)
    env.reset()
    r1 = env.step(np.array([2]))[1][0]  # (2,1) new
    env.step(np.array([0])); env.step(np.array([0]))  # turn around (no new cell)
    r2 = env.step(np.array([2]))[1][0]  # back to (1,1): start cell already visited
    assert r1 == np.float32(0.01) and r2 == 0


# ---- Python literal oracle vs C oracle on objects the reference never uses (doors, keys, lava) ---
def test_c_oracle_matches_python_oracle_full_object_set():
    rng = np.random.default_rng(7)

    class Scratch(mg.MiniGridEnv):
        def __init__(self):
            super().__init__(mission_space=mg.MissionSpace(lambda: "x"), grid_size=9, max_steps=60)

        def _gen_grid(self, w, h):
            self.grid = mg.Grid(w, h)
            self.grid.wall_rect(0, 0, w, h)
            self.grid.set(4, 1, mg.Door("yellow", is_locked=True))
            self.grid.set(4, 2, mg.Door("blue"))
            self.grid.set(4, 3, mg.Wall())
            self.grid.set(2, 2, mg.Key("yellow"))
            self.grid.set(2, 4, mg.Ball("red"))
            self.grid.set(3, 5, mg.Box("purple"))
            self.grid.set(6, 6, mg.Lava())
            self.grid.set(5, 5, mg.Floor())
            self.grid.set(7, 7, mg.Goal())
            self.agent_pos, self.agent_dir = (1, 1), 0

    penv = Scratch()
    for episode in range(6):
        pobs, _ = penv.reset(seed=episode)
        enc = penv.grid.encode()[None]
        cenv = fast.OracleVecEnv(1, enc, np.array([[1, 1, 0]], np.int32), max_steps=60, n_actions=7,
                                 auto_reset=False)
        rgb, sym = cenv.reset()
        assert np.array_equal(sym[0], pobs["image"])
        assert np.array_equal(rgb[0], penv.get_frame(tile_size=8, agent_pov=True))
        for t in range(60):
            a = int(rng.choice([0, 1, 2, 2, 2, 3, 4, 5, 6]))
            pobs, pr, pte, ptr, _ = penv.step(a)
            rgb, r, te, tr, info = cenv.step(np.array([a]))
            assert np.array_equal(info["obs_symbolic"][0], pobs["image"]), (episode, t)
            assert np.array_equal(rgb[0], penv.get_frame(tile_size=8, agent_pov=True)), (episode, t)
            assert (bool(te[0]), bool(tr[0])) == (pte, ptr) and r[0] == np.float32(pr)
            assert (cenv.ax[0], cenv.ay[0], cenv.adir[0]) == (penv.agent_pos[0], penv.agent_pos[1], penv.agent_dir)
            if pte or ptr:
                break
