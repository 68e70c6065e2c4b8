"""GPU tier: the reference-facing Python surface (`src.*`) running on the CUDA path, checked against the golden
fixtures produced by the reference's own code and against the oracle.

  * `ScenarioCreator.create_env` wrapper stack replays the reference traces bit-exactly FROM THE SEED (layout
    generation + step + observation), including StuckPenaltyWrapper.
  * BASELINE config 1: the first rollout of the reference's own PPO (seed 777) is replayed through the CUDA env with
    the recorded actions -- rewards, dones and every 56x56x3 frame (CRC) are identical; the GAE kernel reproduces the
    reference's advantages/returns.
  * batched PPO / FOMAML: rollouts stored on the device are self-consistent with the oracle env driven by the stored
    actions; updates run and change the weights; CUDA-graph rollouts equal eager rollouts in distribution checks.
"""
from __future__ import annotations

import zlib

import numpy as np
import pytest
import torch

import helpers
from oracle import fast

pytestmark = pytest.mark.gpu


def _sc():
    from src.scenario_creator.scenario_creator import ScenarioCreator
    return ScenarioCreator()


def _crc(a):
    return zlib.crc32(np.ascontiguousarray(a).tobytes())


# ---- single-env wrapper stack, from the seed --------------------------------------------------------------
@pytest.mark.parametrize("name", helpers.trace_names())
def test_create_env_replays_reference_trace_from_seed(name):
    from src.wrappers.stuck_penalty_wrapper import StuckPenaltyWrapper
    tr = helpers.load(f"trace_{name}.npz")
    sc = _sc()
    diff, size = str(tr["difficulty"]), int(tr["size"])
    sc.config["difficulties"][diff]["params"]["size"] = size
    env = sc.create_env(diff)
    if bool(tr["stuck_wrapper"]):
        env = StuckPenaltyWrapper(env)
    obs, info = env.reset(seed=int(tr["seed"]))
    u = env.unwrapped
    if int(tr["max_steps"]) != u.max_steps:
        pytest.skip("fixture overrides max_steps after construction")
    assert obs.dtype == np.uint8 and obs.shape == (56, 56, 3) and info == {}
    assert np.array_equal(obs, tr["reset_obs_rgb"][0])
    assert np.array_equal(u.grid.encode(), tr["ep_enc"][0])
    assert (u.agent_pos[0], u.agent_pos[1], u.agent_dir) == tuple(tr["ep_agent"][0])
    for t, a in enumerate(tr["action"]):
        obs, r, te, tr_, info = env.step(int(a))
        assert isinstance(r, float) and isinstance(te, bool) and isinstance(tr_, bool)
        assert np.array_equal(obs, tr["obs_rgb"][t]), t
        assert np.float32(r) == np.float32(tr["reward"][t]), t
        assert (te, tr_) == (bool(tr["terminated"][t]), bool(tr["truncated"][t])), t
        assert (u.agent_pos[0], u.agent_pos[1], u.agent_dir, u.step_count) == tuple(tr["pose"][t]), t
        if bool(tr["stuck_wrapper"]):
            assert info["stuck"] == bool(tr["stuck"][t]), t
        if te or tr_:
            ep = int(tr["episode"][t]) + 1
            obs, _ = env.reset()  # continues the env's RNG stream: next layout of the reference run
            assert np.array_equal(obs, tr["reset_obs_rgb"][ep]), t
            assert np.array_equal(u.grid.encode(), tr["ep_enc"][ep]), t
    env.close()


def test_single_env_surface_details():
    sc = _sc()
    env = sc.create_env("mediumhard", seed=3)
    o1, _ = env.reset(seed=11)
    o2, _ = env.reset(seed=11)  # FOMAML re-seeds at every reset (src/fomaml.py:63,92): same task again
    assert np.array_equal(o1, o2)
    with pytest.raises(ValueError, match="Unknown action"):
        env.unwrapped.step(9)
    frame = env.unwrapped.get_frame()  # fomaml_train.py:107
    assert frame.shape == (16 * 32, 16 * 32, 3) and frame.dtype == np.uint8
    pov = env.unwrapped.get_frame(tile_size=8, agent_pov=True)
    assert np.array_equal(pov, o2)
    # full frame against the literal renderer of the oracle
    from oracle import merlin_ref as mr
    ref = mr.make_env("mediumhard", size=16)
    ref.reset(seed=11)
    assert np.array_equal(frame, ref.unwrapped.get_frame(True, 32, False))
    env.step(1)
    ref.step(1)
    assert np.array_equal(env.unwrapped.get_frame(), ref.unwrapped.get_frame(True, 32, False))
    env.close()


# ---- BASELINE config 1: the reference's own PPO rollout ------------------------------------------------------
def test_config1_reference_ppo_rollout_replayed_on_gpu():
    from merlin_b200 import BatchedMerlinEnv, gae
    fx = helpers.load("ppo_config1_rollout.npz")
    k0 = int(fx["resets_before_rollout"])  # reset(seed) and PPO.__init__'s reset happen before the rollout
    env = BatchedMerlinEnv(1, enc=fx["ep_enc"], agent=fx["ep_agent"], device="cuda:0", auto_reset=False,
                           reset_mode="next")
    for k in range(k0 + 1):  # ... and collect_rollouts() starts with its own reset (src/ppo.py:65)
        obs, _ = env.reset()
        assert _crc(obs[0].cpu().numpy()) == int(fx["reset_crc"][k])
    assert _crc(obs[0].cpu().numpy()) == int(fx["first_state_crc"])
    ep, rewards, dones, ep_ret, ep_len, rets, lens = k0, [], [], 0.0, 0, [], []
    for t, a in enumerate(fx["action"]):
        obs, r, te, tr, _ = env.step(torch.tensor([int(a)], device="cuda:0"))
        assert _crc(obs[0].cpu().numpy()) == int(fx["step_crc"][t]), t
        done = bool(te[0]) or bool(tr[0])
        rewards.append(float(r[0]))
        dones.append(float(done))
        ep_ret += float(r[0])
        ep_len += 1
        if done:
            rets.append(ep_ret)
            lens.append(ep_len)
            ep_ret, ep_len = 0.0, 0
            ep += 1
            obs, _ = env.reset()
            assert _crc(obs[0].cpu().numpy()) == int(fx["reset_crc"][ep]), t
    assert np.array_equal(np.float32(rewards), fx["reward"]) and np.array_equal(np.float32(dones), fx["done"])
    assert lens == fx["episode_lengths"].tolist() and np.allclose(rets, fx["episode_returns"], atol=1e-6)
    dev = "cuda:0"
    adv, ret = gae(torch.tensor(fx["reward"], device=dev), torch.tensor(fx["value"], device=dev),
                   torch.tensor(fx["done"], device=dev), float(fx["last_value"]), float(fx["gamma"]), float(fx["lam"]))
    assert np.array_equal(adv.cpu().numpy(), fx["adv"]) and np.array_equal(ret.cpu().numpy(), fx["ret"])


# ---- batched PPO ----------------------------------------------------------------------------------------------
@pytest.mark.parametrize("graph", [False, True])
def test_batched_ppo_rollout_is_consistent_with_oracle_and_updates(graph):
    from merlin_b200 import codes, layouts
    from src.ppo import PPO
    torch.manual_seed(5)
    N, T, L = 64, 12, 256
    sc = _sc()
    env = sc.create_batched_env("mediumhard", N, device="cuda:0", seeds=range(500, 500 + L), max_steps=9)
    agent = PPO(env, batch_size=N * T, minibatch_size=256, update_epochs=2, ent_coef=0.05, use_cuda_graph=graph)
    assert agent.buffer.states.shape == (T, N, 56, 56, 3) and agent.buffer.states.dtype == torch.uint8
    before = [p.detach().clone() for p in agent.ac.parameters()]
    cells, ag = layouts.generate("mediumhard", 16, range(500, 500 + L))
    ref = fast.OracleVecEnv(N, codes.unpack_to_encoding(cells, 16, 16), ag, max_steps=9)
    n_eps = 0
    for it in range(2):  # with graph=True the second call replays the captured graph
        last_value = agent.collect_rollouts()
        assert last_value.shape == (N,)
        states, actions, logp, rewards, values, dones = agent.buffer.get()
        assert torch.isfinite(logp).all() and torch.isfinite(values).all()
        nxt = torch.cat([states[1:], agent._last_obs.unsqueeze(0)]).cpu().numpy()
        # every rollout starts from a fresh reset of all envs on their next layouts (src/ppo.py:65); the oracle
        # follows the same cursor rule, so it can be driven with the stored actions across both rollouts
        robs, _ = ref.reset()
        assert np.array_equal(states[0].cpu().numpy(), robs), it
        acts = actions.cpu().numpy()
        for t in range(T):
            robs, rr, rte, rtr, rinfo = ref.step(acts[t])
            assert np.array_equal(rewards[t].cpu().numpy(), rr), (it, t)
            assert np.array_equal(dones[t].cpu().numpy(), (rte | rtr).astype(np.float32)), (it, t)
            assert np.array_equal(nxt[t], robs), (it, t)
            n_eps += int((rinfo["episode_length"] > 0).sum())
        assert int(dones.sum()) >= N  # max_steps 9 < T: every env truncates at least once
    assert len(agent.episode_lengths) == len(agent.episode_returns) == n_eps
    assert all(0 < l <= 9 for l in agent.episode_lengths)
    adv, ret = agent.compute_gae(rewards, values, dones, last_value)
    radv, rret = fast.gae(rewards.cpu().numpy(), values.cpu().numpy(), dones.cpu().numpy(), last_value.cpu().numpy(),
                          0.99, 0.95)
    assert np.array_equal(adv.cpu().numpy(), radv) and np.array_equal(ret.cpu().numpy(), rret)
    m = agent.update(last_value)
    assert set(m) == {"pi_loss", "v_loss", "entropy", "kl", "clipfrac", "gradnorm"}
    assert all(np.isfinite(v) for v in m.values()) and 0.5 < m["entropy"] <= np.log(3) + 1e-5
    assert any(not torch.equal(a, b.detach()) for a, b in zip(before, agent.ac.parameters()))


@pytest.mark.parametrize("graph", [False, True], ids=["eager", "cuda_graph"])
def test_batched_ppo_carries_unfinished_episodes_across_rollouts(graph):
    """horizon < max_steps: a rollout continues the episodes the previous one left unfinished (carry_episodes, automatic
    then) -- its first observation is the one the last rollout ended on, episode clocks run on across rollouts, episodes
    longer than the horizon reach the log; with carry_episodes=False every rollout starts from a fresh reset()."""
    from src.ppo import PPO
    N, T, M = 64, 8, 20
    torch.manual_seed(3)
    env = _sc().create_batched_env("mediumhard", N, device="cuda:0", seeds=range(128), max_steps=M)
    agent = PPO(env, batch_size=N * T, minibatch_size=128, update_epochs=1, use_cuda_graph=graph)
    assert agent.carry_episodes
    agent.collect_rollouts()
    sc1 = env.state_numpy()["step_count"]
    assert agent.unfinished_episodes == int((sc1 > 0).sum()) > 0 and sc1.max() == T
    ended_on = agent._last_obs.clone()
    agent.collect_rollouts()
    assert torch.equal(agent.buffer.get()[0][0], ended_on)          # rollout 2 opens where rollout 1 stopped
    sc2 = env.state_numpy()["step_count"]
    assert sc2.max() == 2 * T and int((sc2 == 2 * T).sum()) >= int((sc1 == T).sum()) // 2
    agent.collect_rollouts()                                         # 24 steps > max_steps 20: truncations at 20
    assert max(agent.episode_lengths) == M > T
    assert env.state_numpy()["step_count"].max() < M
    torch.manual_seed(3)
    env2 = _sc().create_batched_env("mediumhard", N, device="cuda:0", seeds=range(128), max_steps=M)
    fresh = PPO(env2, batch_size=N * T, minibatch_size=128, update_epochs=1, use_cuda_graph=graph, carry_episodes=False)
    for _ in range(3):
        fresh.collect_rollouts()
        assert env2.state_numpy()["step_count"].max() == T         # every rollout restarted all envs
    assert all(l <= T for l in fresh.episode_lengths)


def test_batched_ppo_stored_logp_and_values_match_evaluate():
    from src.ppo import PPO
    torch.manual_seed(1)
    env = _sc().create_batched_env("medium", 32, device="cuda:0", seeds=range(64))
    agent = PPO(env, batch_size=32 * 4, minibatch_size=64, update_epochs=1)
    agent.collect_rollouts()
    s, a, logp, r, v, d = agent.buffer.get()
    with torch.no_grad():
        lp, _, val = agent.ac.evaluate(s.reshape(-1, 56, 56, 3), a.reshape(-1))
    assert torch.allclose(lp.view_as(logp), logp, atol=1e-5) and torch.allclose(val.view_as(v), v, atol=1e-4)
    agent.train(total_steps=2 * 32 * 4)  # reference PPO.train loop


def test_ppo_symbolic_rollout_storage_equals_frame_storage():
    """obs_storage='symbolic' keeps 147 B per step and renders minibatches on read: same rollout, same update."""
    from src.actor_critic import space_to_depth4
    from src.ppo import PPO
    runs = {}
    for mode in ("rgb", "symbolic", "symbolic_u8"):
        torch.manual_seed(9)
        env = _sc().create_batched_env("hard", 48, device="cuda:0", seeds=range(96), max_steps=20, want_symbolic=True)
        agent = PPO(env, batch_size=48 * 10, minibatch_size=96, update_epochs=2, ent_coef=0.05, obs_storage=mode[:8],
                    minibatch_frames=torch.uint8 if mode.endswith("u8") else torch.float32)
        lv = agent.collect_rollouts()
        states, actions, logp, rewards, values, dones = agent.buffer.get()
        frames = states if mode == "rgb" else env.render(states.reshape(-1, 7, 7, 3)).reshape(10, 48, 56, 56, 3)
        if mode != "rgb":
            assert states.shape == (10, 48, 7, 7, 3)
            blk = env.render(states.reshape(-1, 7, 7, 3), blocked=True)
            ref_blk = space_to_depth4(frames.reshape(-1, 56, 56, 3)).permute(0, 2, 3, 1).to(torch.uint8)
            assert torch.equal(blk, ref_blk)
        metrics = agent.update(lv)
        runs[mode] = (frames.clone(), actions.clone(), rewards.clone(), values.clone(), lv.clone(), metrics,
                      [p.detach().clone() for p in agent.ac.parameters()])
    # minibatches written by the render kernel as float32 or as uint8 + a PyTorch cast: the same values, the same update
    for ma, mb_ in zip(runs["symbolic"][6], runs["symbolic_u8"][6]):
        assert torch.allclose(ma, mb_, atol=1e-6)  # (cuDNN's backward kernels are free to reorder their sums)
    for k, v in runs["symbolic"][5].items():
        assert abs(v - runs["symbolic_u8"][5][k]) < 1e-5 * max(1.0, abs(v)), k
    a, b = runs["rgb"], runs["symbolic"]
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    assert torch.allclose(a[3], b[3], atol=1e-6) and torch.allclose(a[4], b[4], atol=1e-6)
    for k in a[5]:
        assert abs(a[5][k] - b[5][k]) < 1e-4 * max(1.0, abs(a[5][k])), (k, a[5][k], b[5][k])
    for pa, pb in zip(a[6], b[6]):
        assert torch.allclose(pa, pb, atol=1e-5)


def test_batched_ppo_with_a_single_env_uses_reference_shaped_buffers():
    """N = 1 batched env = the reference regime (1 env x T steps) with device-resident rollouts: [T] buffers."""
    from src.ppo import PPO
    torch.manual_seed(0)
    env = _sc().create_batched_env("mediumhard", 1, device="cuda:0", seeds=range(300, 364), max_steps=30)
    agent = PPO(env, batch_size=96, minibatch_size=32, update_epochs=2)
    lv = agent.collect_rollouts()
    assert agent.buffer.rewards.shape == (96,) and agent.buffer.states.shape == (96, 56, 56, 3) and lv.shape == (1,)
    assert int(agent.buffer.dones.sum()) >= 3 and len(agent.episode_lengths) == int(agent.buffer.dones.sum())
    m = agent.update(lv)
    assert np.isfinite(m["pi_loss"]) and np.isfinite(m["kl"])


def test_single_env_ppo_reference_loop_runs_on_the_cuda_env():
    from src.ppo import PPO
    torch.manual_seed(0)
    env = _sc().create_env("medium")
    env.reset(seed=4)
    agent = PPO(env, batch_size=64, minibatch_size=32, update_epochs=1, device="cuda:0")
    assert agent.buffer.states.shape == (64, 56, 56, 3) and agent.buffer.states.dtype == torch.float32
    lv = agent.collect_rollouts()
    assert isinstance(lv, float)
    m = agent.update(lv)
    assert np.isfinite(m["pi_loss"]) and np.isfinite(m["gradnorm"])
    adv, ret = agent.compute_gae(agent.buffer.rewards, agent.buffer.values, agent.buffer.dones, lv)
    radv, rret = fast.gae(agent.buffer.rewards.cpu().numpy()[:, None], agent.buffer.values.cpu().numpy()[:, None],
                          agent.buffer.dones.cpu().numpy()[:, None], np.float32([lv]), 0.99, 0.95)
    assert np.allclose(adv.cpu().numpy(), radv[:, 0], rtol=1e-6, atol=1e-7)
    assert np.allclose(ret.cpu().numpy(), rret[:, 0], rtol=1e-6, atol=1e-7)
    env.close()


# ---- FOMAML ---------------------------------------------------------------------------------------------------
def test_fomaml_task_batched_meta_step_matches_per_task_loop():
    from merlin_b200 import codes, layouts
    from src.fomaml import FOMAML
    torch.manual_seed(3)
    sc = _sc()
    fo = FOMAML(sc, lr_inner=0.01, lr_outer=3e-4, device="cuda:0", difficulty="mediumhard")
    assert (fo.gamma, fo.lam, fo.vf_coef, fo.ent_coef, fo.clip_eps) == (0.995, 0.95, 0.5, 0.05, 0.2)
    seeds = [7, 11, 100003, 5]
    B, k = len(seeds), 24
    # (a) trajectories: task b == env b, restarted on its own layout; frames/rewards reproducible by the oracle
    env = fo._task_env(seeds)
    traj = fo.collect_trajectory(env, fo.meta_policy, steps=k)
    assert traj["obs"].shape == (k, B, 56, 56, 3) and traj["act"].shape == (k, B) and traj["last_val"].shape == (B,)
    cells, ag = layouts.generate("mediumhard", 16, seeds)
    ref = fast.OracleVecEnv(B, codes.unpack_to_encoding(cells, 16, 16), ag, reset_mode="same")
    robs, _ = ref.reset()
    assert np.array_equal(traj["obs"][0].cpu().numpy(), robs)
    for t in range(k):
        robs, rr, rte, rtr, _ = ref.step(traj["act"][t].cpu().numpy())
        assert np.array_equal(traj["rew"][t].cpu().numpy(), rr)
        assert np.array_equal(traj["done"][t].cpu().numpy(), (rte | rtr).astype(np.float32))
        if t + 1 < k:
            assert np.array_equal(traj["obs"][t + 1].cpu().numpy(), robs)
    # (b) task-batched loss under stacked weights == the reference-style per-task loss on each column
    meta = fo.meta_policy
    fast_w = {n: p.detach().unsqueeze(0).repeat((B,) + (1,) * p.dim()).requires_grad_(True)
              for n, p in meta.named_parameters()}
    loss_b, stats = fo.compute_loss(traj, meta, params=fast_w)
    assert loss_b.shape == (B,)
    for b in range(B):
        single = {"obs": traj["obs"][:, b].float(), "act": traj["act"][:, b], "rew": traj["rew"][:, b],
                  "val": traj["val"][:, b], "logp": traj["logp"][:, b], "done": traj["done"][:, b],
                  "last_val": traj["last_val"][b]}
        l1, s1 = fo.compute_loss(single, meta)
        assert abs(float(l1) - float(loss_b[b])) < 2e-4 * max(1.0, abs(float(l1))), (b, float(l1), float(loss_b[b]))
        # advantages: kernel vs the reference's numpy loop restated in the oracle (normalise, then ret = val + adv)
        from oracle import merlin_ref as mr
        _, adv_ref, ret_ref = mr.gae_fomaml(single["rew"].cpu().numpy(), single["val"].cpu().numpy(),
                                         single["done"].cpu().numpy(), float(single["last_val"]), 0.995, 0.95)
        adv_k, ret_k = fo._advantages(single)
        assert np.allclose(adv_k.cpu().numpy(), adv_ref, rtol=1e-5, atol=1e-6)
        assert np.allclose(ret_k.cpu().numpy(), ret_ref, rtol=1e-5, atol=1e-6)
    # (c) a full meta step moves the meta weights and reports finite numbers
    before = [p.detach().clone() for p in meta.parameters()]
    avg_loss, avg_rew, avg_steps, qstats = fo.meta_train_step(seeds, k_support=k, k_query=k)
    assert np.isfinite(avg_loss) and np.isfinite(avg_rew) and avg_steps > 0
    assert {"pi_loss", "v_loss", "entropy", "kl", "clipfrac", "loss"} <= set(qstats)
    assert any(not torch.equal(a, b.detach()) for a, b in zip(before, meta.parameters()))
    assert all(torch.isfinite(p).all() for p in meta.parameters())


def test_fomaml_reference_default_arguments_run_the_batched_path():
    """FOMAML(scenario_creator) with the reference's defaults (device="cpu", difficulty="medium"): the task-batched
    meta step runs on the GPU the env kernels write to -- no device mismatch between frames and weights."""
    from src.fomaml import FOMAML
    torch.manual_seed(0)
    fo = FOMAML(_sc())
    assert fo.device.type == "cuda" and next(fo.meta_policy.parameters()).device.type == "cuda"
    loss, rew, steps, stats = fo.meta_train_step([3, 4, 5], k_support=12, k_query=12)
    assert np.isfinite(loss) and np.isfinite(stats["kl"])
    r, n, g = fo.few_shot_evaluate([200000, 200001], k_support=8, adapt_steps=1)
    assert r.shape == (2,) and np.all(n >= 1)


def test_fomaml_single_env_reference_path():
    from src.fomaml import FOMAML
    torch.manual_seed(0)
    sc = _sc()
    fo = FOMAML(sc, device="cuda:0", difficulty="medium")
    env = sc.create_env("medium", seed=9)
    batch = fo.collect_trajectory(env, fo.fast_policy, steps=16, task_seed=9)
    assert batch["obs"].shape == (16, 56, 56, 3) and batch["rew"].shape == (16,)
    loss, stats = fo.compute_loss(batch, fo.fast_policy)
    assert loss.dim() == 0 and np.isfinite(float(loss)) and np.isfinite(stats["kl"])
    env.close()


# ---- batched deterministic evaluation ----------------------------------------------------------
def test_batched_evaluation_matches_sequential_reference_loop():
    """evaluate_seeds == the reference's one-seed-at-a-time greedy loop (ppo/ppo_train.py:43-69) run on the oracle env
    with the same policy."""
    from oracle import merlin_ref as mr
    from src.actor_critic import CNNActorCritic
    from src.evaluation import evaluate_seeds
    torch.manual_seed(2)
    policy = CNNActorCritic((56, 56, 3), 3).to("cuda:0")
    seeds = [200000, 200001, 200002, 7, 8]
    ret, length, goal = evaluate_seeds(policy, "medium", 8, seeds, device="cuda:0", poll=16)
    # the same evaluation stepped eagerly (no CUDA graph), and through the frame path (a caller-supplied act_fn)
    ret_e, length_e, goal_e = evaluate_seeds(policy, "medium", 8, seeds, device="cuda:0", poll=16, use_cuda_graph=False)
    ret_f, length_f, goal_f = evaluate_seeds(policy, "medium", 8, seeds, device="cuda:0", poll=16,
                                             act_fn=lambda o: policy.act(o, deterministic=True)[0])
    for other in ((ret_e, length_e, goal_e), (ret_f, length_f, goal_f)):
        assert np.array_equal(other[1], length) and np.allclose(other[0], ret) and np.array_equal(other[2], goal)
    for k, s in enumerate(seeds):
        env = mr.make_env("medium", size=8)
        obs, _ = env.reset(seed=s)
        done, total, steps, te = False, 0.0, 0, False
        while not done:
            with torch.no_grad():
                a = policy.act(torch.as_tensor(obs, device="cuda:0").unsqueeze(0), deterministic=True)[0]
            obs, r, te, tr, _ = env.step(int(a.item()))
            total += r
            steps += 1
            done = te or tr
        assert steps == int(length[k]) and abs(total - ret[k]) < 1e-6 and bool(goal[k]) == bool(te), (s, steps, length[k])


def test_fomaml_few_shot_evaluation_batched():
    """adapt_steps = 0 is zero-shot evaluation of the meta weights; with adaptation every task gets its own weights."""
    from src.evaluation import evaluate_seeds
    from src.fomaml import FOMAML
    torch.manual_seed(4)
    fo = FOMAML(_sc(), lr_inner=0.05, device="cuda:0", difficulty="medium")
    seeds = [300000, 300001, 300002, 300003, 300004, 300005]
    r0, n0, g0 = fo.few_shot_evaluate(seeds, k_support=32, adapt_steps=0)
    r1, n1, g1 = evaluate_seeds(fo.meta_policy, "medium", 16, seeds, device="cuda:0")
    assert np.array_equal(n0, n1) and np.allclose(r0, r1) and np.array_equal(g0, g1)
    before = [p.detach().clone() for p in fo.meta_policy.parameters()]
    r2, n2, g2 = fo.few_shot_evaluate(seeds, k_support=32, adapt_steps=2)
    assert r2.shape == (6,) and np.all(n2 >= 1) and np.all(n2 <= 1024) and np.all((r2 > 0) == g2)
    assert all(torch.equal(a, b) for a, b in zip(before, fo.meta_policy.parameters()))  # the meta weights are untouched


# ---- symbolic-only mode (automatic kernel choice: no frames requested -> state-phase-only kernel) ----------------
def _to_np(x):
    return x.detach().cpu().numpy()


@pytest.mark.parametrize("sym_kernel", [False, True], ids=["auto_kernel", "lane_per_env_kernel"])
@pytest.mark.parametrize("N,n_actions", [(1, 3), (37, 3), (2048, 3), (5000, 3), (70000, 3), (300, 7)])
def test_symbolic_only_observations(N, n_actions, sym_kernel):
    """want_rgb=False: no frames are written -- the state-phase-only kernel (lane per env), or for batches of up to 2048
    envs the warp-per-env kernel (automatic choice); both checked: symbolic images, rewards, flags, poses."""
    from merlin_b200 import BatchedMerlinEnv, codes, layouts
    from test_gpu_parity import _object_layouts
    rng = np.random.default_rng(N)
    if n_actions == 3:
        cells, agent = layouts.generate("hard", 16, range(200))
        enc = codes.unpack_to_encoding(cells, 16, 16)
    else:
        enc, agent = _object_layouts(rng, 64, 11)
    env = BatchedMerlinEnv(N, enc=enc, agent=agent, device="cuda:0", n_actions=n_actions, max_steps=13, want_rgb=False)
    assert env.obs is None and ("warp" if N <= 2048 else "sym") in env.step_kernel()
    if sym_kernel:
        env.set_kernel_choice(5)
        assert "sym" in env.step_kernel()
    ref = fast.OracleVecEnv(N, enc, agent, n_actions=n_actions, max_steps=13, want_rgb=False)
    rgb, sym = env.reset()
    assert rgb is None and np.array_equal(_to_np(sym), ref.reset()[1])
    for t in range(40):
        a = rng.integers(0, n_actions, N)
        obs, r, te, tr, info = env.step(torch.as_tensor(a, device="cuda:0"))
        _, rr, rte, rtr, rinfo = ref.step(a)
        assert obs is None
        assert np.array_equal(_to_np(info["obs_symbolic"]), rinfo["obs_symbolic"]), t
        assert np.array_equal(_to_np(r), rr) and np.array_equal(_to_np(te), rte) and np.array_equal(_to_np(tr), rtr), t
        assert np.array_equal(_to_np(info["episode_length"]), rinfo["episode_length"]), t
    assert np.array_equal(env.pose_numpy(), helpers.get_pose(ref))


# ---- fully observable observations (scenario.yaml: observation.fully_observable / flatten) ----------------------------
def test_fully_observable_and_flatten_match_the_reference_wrappers():
    from oracle import merlin_ref as mr
    from merlin_b200 import BatchedMerlinEnv, codes, layouts
    sc = _sc()
    sc.obs_cfg = {"fully_observable": True, "flatten": True}
    env = sc.create_env("hardest")
    ref = mr.make_env("hardest", size=16, fully_observable=True, flatten=True)
    obs, _ = env.reset(seed=21)
    robs, _ = ref.reset(seed=21)
    assert obs.shape == (16 * 16 * 3,) and obs.dtype == np.uint8 and env.observation_space.shape == (768,)
    assert np.array_equal(obs, robs)
    rng = np.random.default_rng(0)
    for t in range(60):
        a = int(rng.integers(0, 3))
        obs, r, te, tr, _ = env.step(a)
        robs, rr, rte, rtr, _ = ref.step(a)
        assert np.array_equal(obs, robs) and np.float32(r) == np.float32(rr) and (te, tr) == (rte, rtr), t
        if te or tr:
            obs, _ = env.reset()
            assert np.array_equal(obs, ref.reset()[0])
    env.close()
    # batched, 7 actions with doors / keys / carried objects: every env's full grid incl. the agent marker
    from test_gpu_parity import _object_layouts
    enc, agent = _object_layouts(np.random.default_rng(5), 32, 11)
    N = 500
    venv = BatchedMerlinEnv(N, enc=enc, agent=agent, device="cuda:0", n_actions=7, max_steps=30)
    oref = fast.OracleVecEnv(N, enc, agent, n_actions=7, max_steps=30)
    venv.reset()
    oref.reset()
    for t in range(25):
        a = rng.integers(0, 7, N)
        venv.step(torch.as_tensor(a, device="cuda:0"))
        oref.step(a)
    full = venv.full_observation().cpu().numpy()
    want = np.stack([oref.gt, oref.gc, oref.gs], axis=-1).reshape(N, 11, 11, 3).transpose(0, 2, 1, 3).copy()
    want[np.arange(N), oref.ax, oref.ay] = np.stack([np.full(N, 10), np.zeros(N, int), oref.adir], axis=-1)
    assert full.shape == (N, 11, 11, 3) and np.array_equal(full, want)


def test_ppo_mlp_on_flattened_fully_observable_env():
    """The reference's MLP branch (src/ppo.py:38-41): 1-D observations -> MLPActorCritic."""
    from src.ppo import PPO
    torch.manual_seed(0)
    sc = _sc()
    sc.obs_cfg = {"fully_observable": True, "flatten": True}
    env = sc.create_env("medium")
    env.reset(seed=2)
    agent = PPO(env, batch_size=48, minibatch_size=24, update_epochs=1, device="cuda:0")
    assert not agent.use_cnn and agent.obs_shape == (768,)
    m = agent.update(agent.collect_rollouts())
    assert np.isfinite(m["pi_loss"]) and np.isfinite(m["entropy"])
    env.close()
