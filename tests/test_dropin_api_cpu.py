"""CPU tier: the drop-in Python surface of the reference (`src.*`) -- names, signatures, error behaviour and the
host-side logic that needs no device (YAML table, registry, action map, rollout storage, rank sharding, the
stacked-weight policy evaluation FOMAML uses).  Compute entry points must refuse to run without CUDA."""
from __future__ import annotations

import inspect

import pytest
import torch

import src  # noqa: F401  (product mirror on sys.path via conftest)
from src import CNNActorCritic, MLPActorCritic, RolloutBuffer, get_device, layer_init  # reference src/__init__.py:1-4
from src.fomaml import FOMAML, _clip_coef, _logits_value
from src.ppo import PPO
from src.scenario_creator.scenario_creator import ScenarioCreator
from src.wrappers.stuck_penalty_wrapper import StuckPenaltyWrapper
from src.wrappers.three_action_wrapper import ThreeActionWrapper


def _params(fn):
    return [p for p in inspect.signature(fn).parameters if p != "self"]


def test_reference_signatures_are_kept():
    # reference src/ppo.py:10-23, src/fomaml.py:9-16,54,110,158, src/rollout_buffer.py:4,15,24,
    # src/scenario_creator/scenario_creator.py:11,35,59-73, src/wrappers/*.py
    assert _params(PPO.__init__)[:11] == ["env", "lr", "gamma", "lam", "clip_eps", "update_epochs", "batch_size",
                                          "minibatch_size", "vf_coef", "ent_coef", "device"]
    d = {k: v.default for k, v in inspect.signature(PPO.__init__).parameters.items()}
    assert (d["lr"], d["gamma"], d["lam"], d["clip_eps"], d["update_epochs"], d["batch_size"], d["minibatch_size"],
            d["vf_coef"], d["ent_coef"], d["device"]) == (3e-4, 0.99, 0.95, 0.2, 10, 2048, 256, 0.5, 0.01, "cpu")
    assert _params(PPO.compute_gae) == ["rewards", "values", "dones", "last_value"]
    assert _params(PPO.update) == ["last_value"] and _params(PPO.train) == ["total_steps"]
    assert hasattr(PPO, "collect_rollouts") and hasattr(PPO, "_obs_to_tensor")
    assert _params(FOMAML.__init__)[:5] == ["scenario_creator", "lr_inner", "lr_outer", "device", "difficulty"]
    assert _params(FOMAML.collect_trajectory)[:4] == ["env", "policy", "steps", "task_seed"]
    assert _params(FOMAML.compute_loss)[:2] == ["batch", "policy"]
    assert _params(FOMAML.meta_train_step) == ["task_seeds", "k_support", "k_query"]
    assert _params(RolloutBuffer.__init__)[:4] == ["buffer_size", "obs_shape", "device", "is_discrete"]
    assert _params(RolloutBuffer.add) == ["state", "action", "logprob", "value", "reward", "done"]
    assert _params(ScenarioCreator.__init__) == ["config_path"]
    assert _params(ScenarioCreator.create_env) == ["difficulty", "seed"]
    assert _params(StuckPenaltyWrapper.__init__) == ["env", "max_stay", "penalty"]
    assert _params(ThreeActionWrapper.__init__) == ["env"]
    for name in ("sample_scenarios", "get_env_id", "get_logging_params", "get_observation_params", "get_env_size_str",
                 "create_batched_env"):
        assert hasattr(ScenarioCreator, name)


def test_scenario_creator_yaml_and_errors(tmp_path):
    with pytest.raises(FileNotFoundError):
        ScenarioCreator(str(tmp_path / "nope.yaml"))
    sc = ScenarioCreator()
    assert sc.get_env_id("mediumhard") == "MERLIN-MediumHard-v0"
    assert sc.get_env_size_str("hard") == "16x16"
    assert sc.get_observation_params() == {"fully_observable": False, "flatten": False}
    assert sc.seed == 42 and sc.global_cfg == {} and sc.get_logging_params() == {}
    with pytest.raises(ValueError, match="Unknown difficulty"):
        sc.create_env("impossible")
    with pytest.raises(ValueError, match="Unknown difficulty"):
        sc.create_batched_env("impossible", 4)
    bad = tmp_path / "two_sizes.yaml"
    bad.write_text("difficulties:\n  a: {env_id: MERLIN-16x16-v0}\n  b: {env_id: MERLIN-32x32-v0}\n")
    with pytest.raises(ValueError, match="Multiple grid sizes"):
        ScenarioCreator(str(bad))


def test_registry_and_wrapper_stack_without_device():
    import src.custom_envs.register as reg
    assert set(reg.registry) == {"MERLIN-Easy-v0", "MERLIN-Medium-v0", "MERLIN-MediumHard-v0", "MERLIN-Hard-v0",
                                 "MERLIN-Hardest-v0"}
    with pytest.raises(KeyError):
        reg.make("MERLIN-Nope-v0")
    env = ScenarioCreator().create_env("mediumhard", seed=5)
    assert env.action_space.n == 3
    assert env.observation_space.shape == (56, 56, 3)
    u = env.unwrapped
    assert (u.max_steps, u.width, u.height, u.agent_view_size, u.see_through_walls) == (1024, 16, 16, 7, False)
    assert [int(env.action(a)) for a in (0, 1, 2)] == [int(u.actions.left), int(u.actions.right), int(u.actions.forward)]
    assert len(u.actions) == 7 and u.action_space.n == 7
    with pytest.raises(RuntimeError):
        env.step(0)  # before reset
    big = ScenarioCreator()
    big.config["difficulties"]["hard"]["params"]["size"] = 32
    assert big.create_env("hard").unwrapped.max_steps == 4 * 32 * 32  # base_env.py:32-33


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device behaviour")
def test_compute_paths_refuse_to_run_without_cuda():
    env = ScenarioCreator().create_env("easy")
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        env.reset(seed=1)
    with pytest.raises(RuntimeError, match="no CPU fallback|CUDA"):
        ScenarioCreator().create_batched_env("easy", 8, seeds=range(8))
    from merlin_b200 import gae
    with pytest.raises(RuntimeError, match="GPU only"):
        gae(torch.zeros(4), torch.zeros(4), torch.zeros(4), 0.0)


def test_rollout_buffer_reference_semantics():
    buf = RolloutBuffer(4, (2, 2, 3), "cpu", is_discrete=True)
    assert buf.states.shape == (4, 2, 2, 3) and buf.states.dtype == torch.float32
    assert buf.actions.dtype == torch.long and buf.max_size == 4 and buf.ptr == 0
    for k in range(5):  # wraps like the reference ring (src/rollout_buffer.py:15-22)
        buf.add(torch.full((2, 2, 3), float(k)), torch.tensor(k % 3), torch.tensor(-0.1 * k), torch.tensor(0.5 * k),
                torch.tensor(float(k)), torch.tensor(float(k == 2)))
    assert buf.ptr == 1
    states, actions, logp, rew, val, done = buf.get()
    assert buf.ptr == 0
    assert states[0, 0, 0, 0] == 4 and actions.tolist() == [1, 1, 2, 0] and rew.tolist() == [4, 1, 2, 3]
    assert done.tolist() == [0, 0, 1, 0] and val[3] == 1.5 and abs(float(logp[1]) + 0.1) < 1e-7
    cont = RolloutBuffer(4, (3,), "cpu", is_discrete=False)
    assert cont.actions.dtype == torch.float32


def test_rollout_buffer_batched_time_major():
    buf = RolloutBuffer(6 * 5, (56, 56, 3), "cpu", num_envs=5, obs_dtype=torch.uint8)
    assert buf.horizon == 6 and buf.states.shape == (6, 5, 56, 56, 3) and buf.states.dtype == torch.uint8
    slot = buf.obs_slot(2)
    slot.fill_(7)  # what the env kernel does through out_obs
    buf.ptr = 2
    buf.add(None, torch.arange(5), torch.zeros(5), torch.ones(5), torch.full((5,), 2.0), torch.zeros(5))
    assert buf.states[2].min() == 7 and buf.actions[2].tolist() == [0, 1, 2, 3, 4] and buf.rewards[2, 4] == 2
    assert slot.data_ptr() == buf.states[2].data_ptr() and slot.is_contiguous()
    with pytest.raises(ValueError):
        RolloutBuffer(7, (3,), "cpu", num_envs=2)
    one = RolloutBuffer(3, (56, 56, 3), "cpu", num_envs=1, obs_dtype=torch.uint8)
    assert one.obs_slot(1).shape == (1, 56, 56, 3) and one.obs_slot(1).data_ptr() == one.states[1].data_ptr()


def test_actor_critic_state_dict_keys_and_math():
    torch.manual_seed(0)
    ac = CNNActorCritic((56, 56, 3), 3)
    keys = set(ac.state_dict())
    # checkpoint compatibility with reference src/actor_critic.py (module names and Sequential indices)
    for k in ("actor_extractor.network.0.weight", "actor_extractor.network.2.weight", "actor_extractor.network.4.bias",
              "critic_extractor.network.0.weight", "actor.0.weight", "actor.2.bias", "critic.0.weight", "critic.2.weight"):
        assert k in keys, k
    assert sum(p.numel() for p in ac.parameters()) == 744_772
    obs_u8 = torch.randint(0, 256, (5, 56, 56, 3), dtype=torch.uint8)
    a, logp, v = ac.act(obs_u8)
    a2, logp2, v2 = ac.act(obs_u8.float(), deterministic=True)
    assert a.shape == logp.shape == v.shape == (5,) and torch.allclose(v, v2)
    dist = torch.distributions.Categorical(logits=ac(obs_u8)[0])
    lp, ent, val = ac.evaluate(obs_u8, a)
    assert torch.allclose(lp, dist.log_prob(a), atol=1e-6) and torch.allclose(ent, dist.entropy(), atol=1e-6)
    assert torch.equal(a2, ac(obs_u8)[0].argmax(1))
    mlp = MLPActorCritic(12, 3)
    assert mlp.act(torch.zeros(2, 12))[0].shape == (2,)
    assert get_device("cpu").type == "cpu" and callable(layer_init)


def test_stacked_weights_match_per_task_modules():
    """FOMAML's vmap(functional_call) path == running each task's own module; per-task clip == clip_grad_norm_."""
    torch.manual_seed(1)
    pol = CNNActorCritic((56, 56, 3), 3)
    B, k = 3, 4
    mods = [CNNActorCritic((56, 56, 3), 3) for _ in range(B)]
    stacked = {n: torch.stack([dict(m.named_parameters())[n].detach() for m in mods]).requires_grad_(True)
               for n, _ in pol.named_parameters()}
    obs = torch.randint(0, 256, (B, k, 56, 56, 3), dtype=torch.uint8)
    lg, v = torch.func.vmap(lambda p, o: _logits_value(pol, p, o))(stacked, obs)
    loss = (lg.pow(2).sum((1, 2)) + v.pow(2).sum(1))
    grads = torch.autograd.grad(loss.sum(), list(stacked.values()))
    coef = _clip_coef(grads, 0.5)
    for b, m in enumerate(mods):
        l2, v2 = m(obs[b])
        assert torch.allclose(lg[b], l2, atol=1e-5) and torch.allclose(v[b], v2, atol=1e-5)
        m.zero_grad()
        (l2.pow(2).sum() + v2.pow(2).sum()).backward()
        ref = [p.grad.clone() for p in m.parameters()]
        for gs, gr in zip(grads, ref):
            assert torch.allclose(gs[b], gr, rtol=1e-4, atol=1e-5)
        total = torch.nn.utils.clip_grad_norm_(m.parameters(), 0.5)
        assert abs(float(coef[b]) - min(1.0, 0.5 / (float(total) + 1e-6))) < 1e-5


def test_shard_is_balanced_and_covers_everything():
    from src import parallel
    for n in (0, 1, 7, 32, 33):
        for w in (1, 2, 3, 8):
            parts = [parallel.shard(range(n), r, w) for r in range(w)]
            assert sum(parts, []) == list(range(n))
            assert max(map(len, parts)) - min(map(len, parts)) <= 1
    assert parallel.world_size() == 1 and parallel.rank() == 0


def test_eval_sweep_job_plan_is_balanced():
    """config 5 shards (arm, size) jobs over ranks, longest first: every job lands once, the 64x64 jobs on distinct ranks."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("eval_sweep", os.path.join(os.path.dirname(__file__), "..", "tools", "eval_sweep.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    jobs = [(4 * s * s + extra, arm, s) for arm, extra in (("a", 0), ("b", 0), ("c", 320)) for s in (16, 24, 32, 48, 64)]
    plan = mod.assign_jobs(jobs, 8)
    assert sorted(j for part in plan for j in part) == sorted(jobs)
    big = [r for r, part in enumerate(plan) for j in part if j[2] == 64]
    assert len(set(big)) == 3
    loads = [sum(j[0] for j in part) for part in plan]
    assert max(loads) <= 4 * 64 * 64 + 320
    assert mod.assign_jobs(jobs, 1)[0] == sorted(jobs, key=lambda j: -j[0])


def test_metrics_helpers():
    from src.metrics.ppo_metrics import aggregate_ppo_update_metrics, compute_episode_stats
    assert aggregate_ppo_update_metrics(1, 2, 3, 4, 5, 6, 0)["kl"] == 0.0
    assert aggregate_ppo_update_metrics(2, 4, 6, 8, 10, 12, 2) == {"pi_loss": 1, "v_loss": 2, "entropy": 3, "kl": 4,
                                                                    "clipfrac": 5, "gradnorm": 6}
    assert compute_episode_stats([], []) == {"episode_return_mean": 0.0, "episode_length_mean": 0.0}
    assert compute_episode_stats([1.0, 0.0], [10, 30]) == {"episode_return_mean": 0.5, "episode_length_mean": 20.0}


def test_blocked_first_layer_is_the_same_function():
    """conv1 as a 2x2/stride-1 conv over space_to_depth4 input == the literal 8x8/stride-4 conv (same parameters)."""
    torch.manual_seed(4)
    ac = CNNActorCritic((56, 56, 3), 3)
    obs = torch.randint(0, 256, (6, 56, 56, 3), dtype=torch.uint8)
    assert ac.blocked_first_layer
    l1, v1 = ac(obs)
    g1 = torch.autograd.grad(l1.pow(2).sum() + v1.sum(), list(ac.parameters()))
    ac.blocked_first_layer = False
    l2, v2 = ac(obs)
    g2 = torch.autograd.grad(l2.pow(2).sum() + v2.sum(), list(ac.parameters()))
    assert torch.allclose(l1, l2, atol=1e-5) and torch.allclose(v1, v2, atol=1e-5)
    for a, b in zip(g1, g2):
        assert torch.allclose(a, b, rtol=1e-3, atol=1e-5)
    l3, v3 = ac(obs.float())  # float NHWC copies (the reference's buffer dtype) take the same path
    assert torch.allclose(l3, l2, atol=1e-6)


def test_load_policy_current_and_legacy_checkpoints(tmp_path):
    """Reference-format checkpoints load; the older shared-trunk layout is mapped onto both trunks (sweep_checkpoints.py:19-50)."""
    from src.evaluation import load_policy
    torch.manual_seed(3)
    ac = CNNActorCritic((56, 56, 3), 3)
    cur = tmp_path / "current.pth"
    torch.save(ac.state_dict(), cur)
    pol, use_cnn = load_policy(str(cur), device="cpu")
    assert use_cnn and not pol.training
    for a, b in zip(pol.state_dict().values(), ac.state_dict().values()):
        assert torch.equal(a, b)
    legacy = {k.replace("actor_extractor.network", "feature_extractor.conv"): v for k, v in ac.state_dict().items()
              if "critic_extractor" not in k}
    old = tmp_path / "legacy.pth"
    torch.save(legacy, old)
    pol2, _ = load_policy(str(old), device="cpu")
    sd = pol2.state_dict()
    assert torch.equal(sd["actor_extractor.network.0.weight"], ac.state_dict()["actor_extractor.network.0.weight"])
    assert torch.equal(sd["critic_extractor.network.0.weight"], ac.state_dict()["actor_extractor.network.0.weight"])
    assert torch.equal(sd["actor.2.weight"], ac.state_dict()["actor.2.weight"])
    mlp = MLPActorCritic(768, 3)
    m = tmp_path / "mlp.pth"
    torch.save(mlp.state_dict(), m)
    pol3, use_cnn3 = load_policy(str(m), device="cpu", obs_shape=(768,))
    assert not use_cnn3 and torch.equal(pol3.actor[0].weight, mlp.actor[0].weight)
