"""Batched deterministic evaluation -- one env per task seed, all tasks in flight at once.

Batched core of the reference's evaluation loops, which run one seed after another with batch-1 inference:
`evaluate_policy` (ppo/ppo_train.py:43-69: `episodes` seeds `seed + ep`), `evaluate_model`
(src/sweep_checkpoints.py:58-78: seeds 200000..), `collect_zero_shot_metrics`
(ppo/analyze_ppo_distribution.py:70-94) and `evaluate_zero_shot` (src/distribution_over_tasks.py:71-96).
Each seed's layout is what `env.reset(seed=s)` builds; the policy acts greedily (`argmax`); an env's first
episode is recorded when it ends (its return and length come from the step kernel's episode counters) and the
env is ignored afterwards ("freeze after done").  The host looks at the device only every `poll` steps.
"""
from __future__ import annotations

import numpy as np
import torch

from merlin_b200 import BatchedMerlinEnv
from merlin_b200 import layouts as _layouts


@torch.no_grad()
def evaluate_seeds(policy, difficulty, size, seeds, device="cuda", max_steps=None, deterministic=True, poll=64,
                   env=None, act_fn=None, use_cuda_graph=True, params=None):
    """Returns (returns f64[len(seeds)], lengths i64[len(seeds)], reached_goal bool[len(seeds)]).
    A CNN policy is evaluated on the lean path: per step the policy's float32 input is rendered from the env's
    147-byte symbolic image, the two trunks run as one fused network (`RolloutPolicy`; `params` = stacked per-task
    weights evaluates task b under its own weights) and ONE launch takes the argmax, steps every env and keeps the
    first-episode record (`BatchedMerlinEnv.policy_step(greedy, record)`); chunks of `poll` steps are replayed from
    one CUDA graph, no host work inside a chunk.  `act_fn(obs u8[B,56,56,3]) -> actions i64[B]` selects the generic
    path instead (frames + torch-side bookkeeping)."""
    from .actor_critic import CNNActorCritic, RolloutPolicy
    seeds = [int(s) for s in seeds]
    B = len(seeds)
    cells, agent = _layouts.generate(difficulty, size, seeds)
    lean = act_fn is None and isinstance(policy, CNNActorCritic) and policy.blocked_first_layer
    if env is None or env.num_envs != B:
        env = BatchedMerlinEnv(B, cells, agent, width=size, height=size, max_steps=max_steps, device=device,
                               reset_mode="same", want_symbolic=lean, want_rgb=not lean)
    else:
        env.upload_layouts(cells, agent)
        lean = lean and env.obs_symbolic is not None
    if params is not None and not lean:
        raise ValueError("per-task weights are evaluated on the lean CNN path only")
    env.set_cursors(np.arange(B, dtype=np.int32))
    dev = env.device
    rec = {"finished": torch.zeros(B, dtype=torch.bool, device=dev), "first_return": torch.zeros(B, dtype=torch.float32, device=dev),
           "first_length": torch.zeros(B, dtype=torch.int32, device=dev), "first_goal": torch.zeros(B, dtype=torch.bool, device=dev)}
    finished = rec["finished"]
    obs, _ = env.reset(frames=not lean)
    was_training = policy.training
    policy.eval()

    if act_fn is not None:  # generic path: caller-chosen actions, bookkeeping in torch
        ret, length, goal = rec["first_return"], rec["first_length"], rec["first_goal"]

        def advance(obs):
            obs, _, term, _, info = env.step(act_fn(obs))
            first = (info["episode_length"] > 0) & ~finished
            ret.copy_(torch.where(first, info["episode_return"], ret))
            length.copy_(torch.where(first, info["episode_length"], length))
            goal.logical_or_(first & term)
            finished.logical_or_(first)
            return obs
    else:
        A = env.n_actions
        if lean:
            rp = RolloutPolicy(policy, params=params, actor_only=True)   # evaluation never reads the value
            heads = torch.zeros((B, 1, A) if params is not None else (2, B, A), dtype=torch.float32, device=dev)
            logits = heads.view(B, A) if params is not None else heads[0]
            policy_in = torch.empty((B, 14, 14, 48), dtype=torch.float32, device=dev)
        else:
            logits = torch.zeros((B, A), dtype=torch.float32, device=dev)
        io = env.make_policy_io(logits, greedy=deterministic, record=rec)

        def advance(obs):
            if lean:
                rp(env.render(env.obs_symbolic, out=policy_in, blocked=True, dtype=torch.float32), out=heads)
            else:
                logits.copy_(policy(obs if isinstance(policy, CNNActorCritic) else obs.reshape(B, -1))[0])
            return env.policy_step(io, frames=not lean)[0]

    if lean and use_cuda_graph and dev.type == "cuda":
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):  # warm-up outside capture (cuDNN plans, allocator); evaluates, does not step
            rp(policy_in.zero_(), out=heads)
        torch.cuda.current_stream(dev).wait_stream(side)
        chunk = torch.cuda.CUDAGraph()
        with torch.cuda.graph(chunk):
            for _ in range(poll):
                advance(None)
        # every env's first episode ends by max_steps; steps replayed beyond that only touch already finished envs
        for _ in range((env.max_steps + poll - 1) // poll):
            chunk.replay()
            if bool(finished.all()):
                break
    else:
        for t in range(env.max_steps):
            obs = advance(obs)
            if (t + 1) % poll == 0 and bool(finished.all()):
                break
    policy.train(was_training)
    return (rec["first_return"].double().cpu().numpy(), rec["first_length"].long().cpu().numpy(),
            rec["first_goal"].cpu().numpy())


def scenario_geometry(sc, difficulty):
    """(layout-routine name, grid size, max_steps or None) of a scenario-table entry, merged with the YAML's global
    section exactly as `ScenarioCreator.create_env` / `create_batched_env` merge it."""
    from src.custom_envs.register import DIFFICULTY_OF
    entry = sc._difficulty_cfg(difficulty)
    params = sc._env_params(entry)
    return DIFFICULTY_OF[entry["env_id"]], int(params.get("size", 16)), params.get("max_steps")


def evaluate_policy(agent, env_or_creator, episodes=3, seed=None, difficulty=None):
    """Signature of ppo/ppo_train.py:43 -- `(rewards, steps_list)` for seeds `seed + ep`.  `env_or_creator` may be a
    ScenarioCreator (then `difficulty` names the scenario), a BatchedMerlinEnv, or a reference-style single env."""
    base = seed if seed is not None else 0
    seeds = [base + ep for ep in range(episodes)]
    max_steps = None
    if hasattr(env_or_creator, "create_batched_env"):
        diff, size, max_steps = scenario_geometry(env_or_creator, difficulty)
    elif isinstance(env_or_creator, BatchedMerlinEnv):
        env = env_or_creator
        diff = env.difficulty or difficulty
        if diff is None:
            raise ValueError("this BatchedMerlinEnv does not know its layout routine: pass difficulty=... "
                             "(envs made by ScenarioCreator.create_batched_env carry it)")
        if env.width != env.height:
            raise ValueError("evaluation layouts are generated for square grids")
        size, max_steps = env.width, env.max_steps
    else:
        u = env_or_creator.unwrapped
        diff, size, max_steps = u.difficulty, u.size, getattr(u, "max_steps", None)
    r, n, _ = evaluate_seeds(agent.ac, diff, size, seeds, max_steps=max_steps,
                             device=agent.device if agent.device.type == "cuda" else "cuda")
    return r.tolist(), n.tolist()


def evaluate_model(sc, policy, difficulty, seeds, device="cuda"):
    """src/sweep_checkpoints.py:58-78 -- (mean reward, mean steps) over `seeds`."""
    diff, size, max_steps = scenario_geometry(sc, difficulty)
    r, n, _ = evaluate_seeds(policy, diff, size, seeds, device=device, max_steps=max_steps)
    return float(np.mean(r)), float(np.mean(n))


def load_policy(model_path, env=None, device="cuda", obs_shape=(56, 56, 3), act_dim=3):
    """A policy from a reference-format `.pth` state_dict (src/sweep_checkpoints.py:19-50): CNN for image observations,
    MLP for flat ones; checkpoints of the older shared-trunk architecture (keys `feature_extractor.conv.*`) are mapped
    onto both the actor and the critic trunk, the remaining keys loaded non-strictly, as the reference does.
    `env` (optional) supplies observation shape and action count; returns (policy in eval mode, use_cnn)."""
    from .actor_critic import CNNActorCritic, MLPActorCritic
    if env is not None:
        space = getattr(env, "observation_space", None)
        if space is not None and hasattr(space, "shape"):
            obs_shape = tuple(space.shape)
        act_dim = getattr(getattr(env, "action_space", None), "n", act_dim)
    use_cnn = len(obs_shape) == 3
    policy = (CNNActorCritic(obs_shape, act_dim) if use_cnn else MLPActorCritic(int(np.prod(obs_shape)), act_dim)).to(device)
    state = torch.load(model_path, map_location=device, weights_only=True)
    if use_cnn and any("feature_extractor" in k for k in state):
        mapped = {}
        for k, v in state.items():
            if "feature_extractor.conv" in k:
                for trunk in ("actor_extractor.network", "critic_extractor.network"):
                    mapped[k.replace("feature_extractor.conv", trunk)] = v.clone()
            else:
                mapped[k] = v
        policy.load_state_dict(mapped, strict=False)
    else:
        policy.load_state_dict(state)
    policy.eval()
    return policy, use_cnn
