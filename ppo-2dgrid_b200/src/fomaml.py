"""FOMAML with task-batched, device-resident rollouts (reference src/fomaml.py:8-223: same constructor,
methods, attributes, hyper-parameters and loss; the data path is rewired).

The reference adapts one task at a time: a batch-1 policy forward and one Python env step per transition.
Here all tasks of a meta-batch advance together: task b is env b of one BatchedMerlinEnv (`reset_mode="same"`:
a finished episode restarts on the task's own layout, src/fomaml.py:92), the support phase is a single batched
forward of the shared meta-policy per step, and after the inner SGD step the per-task fast weights are a stacked
parameter set evaluated with `torch.func.functional_call` under `vmap` (one grouped convolution for all tasks).
GAE for all tasks is one launch of the CUDA kernel over the `[k, B]` rollout (gamma 0.995, lambda 0.95), followed
by the reference's normalise-then-add returns convention (src/fomaml.py:126-127).  With `torch.distributed`
initialised each rank adapts its shard of `task_seeds`; the accumulated first-order meta-gradient is
all-reduced (SUM) once and divided by the GLOBAL task count (src/fomaml.py:207-209).
"""
from __future__ import annotations

from copy import deepcopy

import numpy as np
import torch
import torch.optim as optim
from torch.func import functional_call, vmap

from merlin_b200 import BatchedMerlinEnv, gae as gae_kernel
from merlin_b200 import layouts as _layouts

from . import parallel
from .actor_critic import CNNActorCritic, MLPActorCritic, RolloutPolicy


class FOMAML:
    def __init__(self, scenario_creator, lr_inner=0.01, lr_outer=3e-4, device="cpu", difficulty="medium", sync_init=True):
        self.sc = scenario_creator
        self.difficulty = difficulty
        self.device = torch.device(device)
        if self.device.type != "cuda" and torch.cuda.is_available():
            # the reference's default is device="cpu"; here meta_train_step / few_shot_evaluate run task-batched on the
            # GPU (there is no CPU env), so the policies live where the env kernels write: the current CUDA device
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.lr_inner = lr_inner

        cfg = self.sc.config["difficulties"].get(difficulty)
        if not cfg:
            raise ValueError(f"Unknown difficulty: {difficulty}")
        self.size = int({**self.sc.global_cfg, **cfg.get("params", {})}.get("size", 16))
        if self.sc.obs_cfg.get("fully_observable", False):
            raise NotImplementedError("task-batched FOMAML consumes the egocentric 56x56x3 frames; the fully observable "
                                      "grid is served on the single-env path only (ScenarioCreator.create_env)")
        if self.sc.obs_cfg.get("flatten", False):
            self.use_cnn = False
            self.meta_policy = MLPActorCritic(56 * 56 * 3, 3).to(self.device)
        else:
            self.use_cnn = True
            self.meta_policy = CNNActorCritic((56, 56, 3), 3).to(self.device)
        if sync_init:  # every rank starts from rank 0's weights (a collective: skip it for rank-local use, e.g. eval jobs)
            parallel.broadcast_parameters(self.meta_policy)
        self.meta_optimizer = optim.Adam(self.meta_policy.parameters(), lr=lr_outer)
        self.fast_policy = deepcopy(self.meta_policy).to(self.device)
        self.fast_policy.train()

        self.gamma = 0.995
        self.lam = 0.95
        self.vf_coef = 0.5
        self.ent_coef = 0.05
        self.clip_eps = 0.2
        self._prefetch = None  # (seeds, thread, result box) of prefetch_tasks
        self._envs = {}  # task-batch size -> cached BatchedMerlinEnv
        self._graphs = {}  # (env, steps, per-task weights?) -> captured rollout
        self._phase_graphs = {}  # (phase, env, k) -> captured loss / gradient pass
        self.use_cuda_graph = self.device.type == "cuda"

    # ---- helpers ------------------------------------------------------------------------------------------
    def _obs_to_tensor(self, state):
        state_t = torch.as_tensor(state, device=self.device).to(torch.float32)
        return state_t.unsqueeze(0) if self.use_cnn else state_t.view(1, -1)

    def _fmt(self, obs):
        return obs if self.use_cnn else obs.reshape(obs.shape[0], -1)

    def prefetch_tasks(self, task_seeds):
        """Start building the layouts of a coming meta-batch on a background thread (the host-side `_gen_grid` of this
        rank's shard of `task_seeds`, ~0.2 ms per task) while the GPU works on the current one; `meta_train_step` /
        `few_shot_evaluate` pick the result up when they are called with the same seeds.  Optional."""
        import threading
        seeds = tuple(int(s) for s in parallel.shard(list(task_seeds)))
        if not seeds or (self._prefetch is not None and self._prefetch[0] == seeds):
            return
        box = {}

        def work():
            box["layouts"] = _layouts.generate(self.sc_difficulty(), self.size, list(seeds))
        th = threading.Thread(target=work, daemon=True)
        th.start()
        self._prefetch = (seeds, th, box)

    def _task_layouts(self, seeds):
        pf = self._prefetch
        if pf is not None and pf[0] == tuple(seeds):
            pf[1].join()
            self._prefetch = None
            if "layouts" in pf[2]:
                return pf[2]["layouts"]
        return _layouts.generate(self.sc_difficulty(), self.size, seeds)

    def _task_env(self, task_seeds):
        """One env per task seed, loaded with the layout `env.reset(seed=s)` would build."""
        seeds = [int(s) for s in task_seeds]
        cells, agent = self._task_layouts(seeds)
        B = len(seeds)
        env = self._envs.get(B)
        if env is None:
            dev = self.device if self.device.type == "cuda" else "cuda"
            env = BatchedMerlinEnv(B, cells, agent, width=self.size, height=self.size, device=dev, reset_mode="same",
                                   want_symbolic=True)
            self._envs[B] = env
        else:
            env.upload_layouts(cells, agent)
        env.set_cursors(np.arange(B, dtype=np.int32))
        return env

    def sc_difficulty(self):
        from src.custom_envs.register import DIFFICULTY_OF
        return DIFFICULTY_OF[self.sc.get_env_id(self.difficulty)]

    # ---- rollouts -----------------------------------------------------------------------------------------
    @torch.no_grad()
    def collect_trajectory(self, env, policy, steps=20, task_seed=None, params=None, params_static=False):
        """`steps` transitions from a fresh reset.  `env`: a BatchedMerlinEnv whose B envs are B tasks (`policy`
        acts for all of them; `params` = stacked per-task weights evaluates task b with its own weights), or a
        reference-style single env (then `task_seed` re-seeds every reset, as in the reference).
        Returns the reference's dict; batched tensors are time-major `[steps, B, ...]` (`obs` is rendered lazily from
        the stored symbolic observations the first time it is read)."""
        if not isinstance(env, BatchedMerlinEnv):
            return self._collect_single(env, policy, steps, task_seed)
        if self.use_cuda_graph and policy is self.meta_policy:
            return self._collect_graphed(env, steps, params, params_static)
        buf = self._rollout_buffers(env, steps)
        self._rollout_body(env, policy, params, steps, buf)
        return self._rollout_result(env, buf, steps)

    def _lean(self, env, policy):
        """The fused rollout path: CNN policy on the env's own symbolic observations (no u8 frame is ever written)."""
        return self.use_cnn and getattr(policy, "blocked_first_layer", False) and env.obs_symbolic is not None

    def _rollout_buffers(self, env, steps):
        B, dev = env.num_envs, env.device
        f32 = dict(dtype=torch.float32, device=dev)
        lean = self._lean(env, self.meta_policy)
        buf = {"act": torch.empty((steps, B), dtype=torch.long, device=dev), "rew": torch.empty((steps, B), **f32),
               "val": torch.empty((steps, B), **f32), "logp": torch.empty((steps, B), **f32),
               "done": torch.empty((steps, B), **f32), "ep_ret": torch.empty((steps, B), **f32),
               "ep_len": torch.empty((steps, B), dtype=torch.int32, device=dev), "last_val": torch.empty(B, **f32),
               "heads": torch.zeros((2 * B, 1, 3), **f32)}
        if lean:
            buf["sym"] = torch.empty((steps + 1, B, 7, 7, 3), dtype=torch.uint8, device=dev)
            buf["policy_in"] = torch.empty((B, 14, 14, 48), **f32)
        else:
            buf["obs"] = torch.empty((steps + 1, B, 56, 56, 3), dtype=torch.uint8, device=dev)
        # the step kernel writes reward / done / episode statistics of step t straight into row t, and -- fused policy
        # transition -- the sampled action, its log-probability and the value as well
        scratch = {k: torch.empty(B, dtype=torch.bool, device=dev) for k in ("terminated", "truncated", "stuck")}
        buf["rows"] = [env.make_step_buffers(reward=buf["rew"][t], done=buf["done"][t], episode_return=buf["ep_ret"][t],
                                             episode_length=buf["ep_len"][t], **scratch) for t in range(steps)]
        return buf

    def _policy_ios(self, env, buf, steps, logits, value):
        return [env.make_policy_io(logits, value, action=buf["act"][t], logprob=buf["logp"][t], value_out=buf["val"][t])
                for t in range(steps)]

    def _rollout_body(self, env, policy, params, steps, buf):
        B = env.num_envs
        lean = "sym" in buf
        if lean:
            # the weights do not change during the rollout: both trunks (of every task) packed once into one fused network
            rp = RolloutPolicy(policy, params=params)
            heads = buf["heads"] if params is not None else buf["heads"].view(2, B, 3)
            hv = heads.view(B, 2, 3) if params is not None else None
            logits, value = (hv[:, 0], hv[:, 1, 0]) if params is not None else (heads[0], heads[1, :, 0])
            ios = self._policy_ios(env, buf, steps, logits, value)
            sym = buf["sym"]
            env.reset(out_symbolic=sym[0], frames=False)
            for t in range(steps):
                rp(env.render(sym[t], out=buf["policy_in"], blocked=True, dtype=torch.float32), out=heads)
                env.policy_step(ios[t], out_symbolic=sym[t + 1], out=buf["rows"][t], frames=False)
            rp(env.render(sym[steps], out=buf["policy_in"], blocked=True, dtype=torch.float32), out=heads)
            buf["last_val"].copy_(value)
            buf["_ios"] = ios  # the C structs must outlive a captured graph
            return
        obs = buf["obs"]
        env.reset(out_obs=obs[0])
        hv = buf["heads"].view(B, 2, 3)
        logits, value = hv[:, 0], hv[:, 1, 0]
        ios = self._policy_ios(env, buf, steps, logits, value)
        for t in range(steps):
            lg, v = self._logits_value(policy, params, obs[t])
            logits.copy_(lg); value.copy_(v)
            env.policy_step(ios[t], out_obs=obs[t + 1], out=buf["rows"][t])
        buf["last_val"].copy_(self._logits_value(policy, params, obs[steps])[1])
        buf["_ios"] = ios

    def _logits_value(self, policy, params, obs):
        """(logits `[B, 3]`, value `[B]`) for one frame per task, shared or per-task weights (generic path)."""
        if params is None:
            return policy(self._fmt(obs))
        logits, value = vmap(lambda p, o: _logits_value(policy, p, o.unsqueeze(0)))(params, self._fmt(obs))
        return logits.squeeze(1), value.squeeze(1)

    @staticmethod
    def _rollout_result(env, buf, steps):
        out = _Trajectory({"act": buf["act"], "rew": buf["rew"], "val": buf["val"], "logp": buf["logp"],
                           "done": buf["done"], "last_val": buf["last_val"]})
        out.stats = (buf["ep_len"], buf["ep_ret"])  # `ep_lens` / `ep_rews` are read back (a sync) only when asked for
        if "sym" in buf:
            out["obs_symbolic"] = buf["sym"][:steps]
            out.env = env
        else:
            out["obs"] = buf["obs"][:steps]
        return out

    def _collect_graphed(self, env, steps, params, params_static=False):
        """The whole k-step rollout (input rendering, fused policy network, fused sample/step/store transition)
        replayed from one CUDA graph per (env, steps, shared | per-task weights).  Per-task weights are copied into the
        graph's static stacked tensors before each replay (the graph re-packs them); the meta-policy's own parameters
        are updated in place by the optimiser, so a graph over them stays valid.  The returned tensors are the graph's
        buffers: valid until the next rollout of that kind."""
        # params_static: `params` are themselves static tensors (outputs of a captured adapt pass): the rollout graph
        # reads them in place, nothing is copied per replay
        ident = id(next(iter(params.values()))) if (params is not None and params_static) else 0
        key = (id(env), steps, params is not None, ident)
        g = self._graphs.get(key)
        if g is None:
            buf = self._rollout_buffers(env, steps)
            if params is None:
                static = None
            elif params_static:
                static = {n: p.detach() for n, p in params.items()}
            else:
                static = {n: torch.empty_like(p) for n, p in params.items()}
                for n in static:
                    static[n].copy_(params[n])
            side = torch.cuda.Stream(env.device)
            side.wait_stream(torch.cuda.current_stream(env.device))
            with torch.cuda.stream(side):  # warm-up outside capture (cuDNN plans, allocator)
                if "sym" in buf:
                    RolloutPolicy(self.meta_policy, params=static)(buf["policy_in"].zero_())
                else:
                    self._logits_value(self.meta_policy, static, buf["obs"][0])
            torch.cuda.current_stream(env.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                self._rollout_body(env, self.meta_policy, static, steps, buf)
            g = self._graphs[key] = (graph, buf, static)
        graph, buf, static = g
        if static is not None and not params_static:
            for n in static:
                static[n].copy_(params[n])
        graph.replay()
        return self._rollout_result(env, buf, steps)

    def _phase(self, key, fn):
        """Run `fn` from a CUDA graph captured on first use (after three eager warm-up runs on a side stream: cuDNN plans,
        autograd, allocator); returns fn's outputs -- the graph's static tensors, rewritten by every replay."""
        g = self._phase_graphs.get(key)
        if g is None:
            cur = torch.cuda.current_stream(self.device)
            side = torch.cuda.Stream(self.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                for _ in range(3):
                    fn()
            cur.wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = fn()
            g = self._phase_graphs[key] = (graph, out)
        g[0].replay()
        return g[1]

    def _act(self, policy, params, obs, **kw):
        """Sampled action, its log-probability and the value for one frame per task (torch-side sampling: the
        generic path for callers outside the fused rollouts)."""
        logits, value = self._logits_value(policy, params, obs)
        logp_all = torch.log_softmax(logits, dim=-1)
        a = torch.multinomial(logp_all.exp(), 1).squeeze(-1)
        return a, logp_all.gather(-1, a.unsqueeze(-1)).squeeze(-1), value

    def _collect_single(self, env, policy, steps, task_seed):
        obs_buf, act_buf, rew_buf, val_buf, logp_buf, done_buf = [], [], [], [], [], []
        ep_lens, ep_rews, cur_len, cur_rew = [], [], 0, 0
        state, _ = env.reset(seed=task_seed)
        for _ in range(steps):
            state_t = self._obs_to_tensor(state)
            action, logp, value = policy.act(state_t, deterministic=False)
            state, reward, terminated, truncated, _ = env.step(action.item())
            done = terminated or truncated
            cur_len += 1
            cur_rew += reward
            obs_buf.append(state_t); act_buf.append(action); rew_buf.append(reward)
            val_buf.append(value); logp_buf.append(logp); done_buf.append(done)
            if done:
                ep_lens.append(cur_len); ep_rews.append(cur_rew)
                cur_len, cur_rew = 0, 0
                state, _ = env.reset(seed=task_seed)
        last_val = policy.act(self._obs_to_tensor(state))[2]
        return {"obs": torch.cat(obs_buf), "act": torch.cat(act_buf),
                "rew": torch.tensor(rew_buf, dtype=torch.float32).to(self.device), "val": torch.cat(val_buf),
                "logp": torch.cat(logp_buf), "done": torch.tensor(done_buf, dtype=torch.float32).to(self.device),
                "last_val": last_val, "ep_lens": ep_lens, "ep_rews": ep_rews}

    # ---- loss ---------------------------------------------------------------------------------------------
    def _advantages(self, batch):
        """GAE kernel over `[k]` or `[k, B]`, then per-task normalisation and `ret = val + adv_norm`."""
        rew, val, done = batch["rew"], batch["val"], batch["done"]
        dev = rew.device
        g = dev if dev.type == "cuda" else torch.device("cuda", torch.cuda.current_device())
        last = batch["last_val"]
        last = last.reshape(-1).to(g) if torch.is_tensor(last) else float(last)
        adv, _ = gae_kernel(rew.to(g), val.to(g), done.to(g), last, self.gamma, self.lam)
        adv = adv.to(dev)
        adv = (adv - adv.mean(0, keepdim=True)) / (adv.std(0, keepdim=True) + 1e-8)
        return adv, (val + adv).detach()

    def compute_loss(self, batch, policy, params=None, want_stats=True):
        """PPO-clip loss of the reference (src/fomaml.py:110-156).  Single trajectory -> (loss, stats) as in the
        reference.  Task-batched trajectory (`[k, B]`) -> (per-task loss `[B]`, stats averaged over tasks); with
        `params` (stacked per-task weights) task b is evaluated under its own weights."""
        total, stat5 = self._loss_terms(batch, policy, params, want_stats)
        if not want_stats:  # internal callers that discard the statistics skip their read-back (a host sync)
            return total, {"loss": total}
        s = stat5.tolist()
        return total, {"loss": total, "pi_loss": s[0], "v_loss": s[1], "entropy": s[2], "kl": s[3], "clipfrac": s[4]}

    def _loss_terms(self, batch, policy, params=None, want_stats=True):
        """The loss without any host synchronisation (capturable in a CUDA graph): (loss, f32[5] = pi_loss, v_loss,
        entropy, kl, clipfrac as a device tensor, or None)."""
        adv, ret = self._advantages(batch)
        old_logp = batch["logp"].detach()
        act = batch["act"]
        if adv.dim() == 1:
            new_logp, entropy, new_vals = policy.evaluate(batch["obs"], act)
        elif params is None:
            k, B = act.shape
            if "obs_symbolic" in batch:  # the first layer's float32 input, rendered from the stored symbolic images
                flat = batch.env.render(batch["obs_symbolic"], blocked=True, dtype=torch.float32)
            else:
                obs = batch["obs"]
                flat = self._fmt(obs.reshape((k * B,) + obs.shape[2:]))
            new_logp, entropy, new_vals = policy.evaluate(flat, act.reshape(-1))
            new_logp, entropy, new_vals = new_logp.view(k, B), entropy.view(k, B), new_vals.view(k, B)
        else:
            if "obs_symbolic" in batch:  # task-major [B, k, 14, 14, 48]: gather + render + cast in one kernel
                k, B = act.shape
                index = (torch.arange(k, device=act.device).unsqueeze(0) * B + torch.arange(B, device=act.device).unsqueeze(1))
                obs_b = batch.env.render(batch["obs_symbolic"], index.reshape(-1), blocked=True,
                                         dtype=torch.float32).view(B, k, 14, 14, 48)
            else:
                obs_b = self._fmt_tasks(batch["obs"])  # [B, k, ...]
            logits, vals = vmap(lambda p, o: _logits_value(policy, p, o))(params, obs_b)
            logp_all = torch.log_softmax(logits, dim=-1)  # [B, k, A]
            entropy = (-(logp_all.exp() * logp_all).sum(-1)).t()
            new_logp = logp_all.gather(-1, act.t().unsqueeze(-1)).squeeze(-1).t()
            new_vals = vals.t()
        ratio = torch.exp(new_logp - old_logp)
        surr = torch.min(ratio * adv, torch.clamp(ratio, 1.0 - self.clip_eps, 1.0 + self.clip_eps) * adv)
        pi_loss = -surr.mean(0)
        v_loss = ((new_vals - ret) ** 2).mean(0)
        ent = entropy.mean(0)
        total = pi_loss + self.vf_coef * v_loss - self.ent_coef * ent
        if not want_stats:
            return total, None
        with torch.no_grad():
            kl = (old_logp - new_logp).mean()
            clipfrac = (torch.abs(ratio - 1.0) > self.clip_eps).float().mean()
            stat5 = torch.stack([pi_loss.mean(), v_loss.mean(), ent.mean(), kl, clipfrac])
        return total, stat5

    def _fmt_tasks(self, obs):
        o = obs.transpose(0, 1)  # [B, k, 56, 56, 3]
        return o if self.use_cnn else o.reshape(o.shape[0], o.shape[1], -1)

    @staticmethod
    def _inner_step(fast, per_task_loss, names, lr):
        """One SGD step per task on stacked weights, each task's gradient clipped to norm 0.5 (src/fomaml.py:179-182)."""
        g = torch.autograd.grad(per_task_loss.sum(), [fast[n] for n in names])
        coef = _clip_coef(g, 0.5)
        B = coef.shape[0]
        with torch.no_grad():
            return {n: (fast[n] - lr * gi * coef.view((B,) + (1,) * (gi.dim() - 1))).requires_grad_(True)
                    for n, gi in zip(names, g)}

    # ---- few-shot evaluation ------------------------------------------------------------------------------
    def few_shot_evaluate(self, task_seeds, k_support=256, adapt_steps=1, lr_inner=None):
        """Adapt-then-evaluate for many tasks at once (batched core of `evaluate_few_shot`,
        src/distribution_over_tasks.py:132-209, and fomaml/analyze_fomaml_distribution.py:54-86): every task starts
        from the meta weights, takes `adapt_steps` inner SGD steps, each on a fresh `k_support`-step rollout of its own
        layout under its current weights, then plays one greedy episode.
        Returns (returns f64[B], lengths i64[B], reached_goal bool[B]) in the order of `task_seeds`."""
        from .evaluation import evaluate_seeds
        seeds = [int(s) for s in task_seeds]
        meta = self.meta_policy
        names = [n for n, _ in meta.named_parameters()]
        lr = self.lr_inner if lr_inner is None else lr_inner
        env = self._task_env(seeds)
        fast = _stack(meta, len(seeds))
        for _ in range(adapt_steps):
            support = self.collect_trajectory(env, meta, steps=k_support, params=fast)
            loss, _ = self.compute_loss(support, meta, params=fast, want_stats=False)
            fast = self._inner_step(fast, loss, names, lr)

        with torch.no_grad():
            return evaluate_seeds(meta, self.sc_difficulty(), self.size, seeds, device=env.device, env=env,
                                  params={n: p.detach() for n, p in fast.items()})

    # ---- meta step ----------------------------------------------------------------------------------------
    def meta_train_step(self, task_seeds, k_support=50, k_query=50):
        task_seeds = list(task_seeds)
        n_global = len(task_seeds)
        my_seeds = parallel.shard(task_seeds)
        meta = self.meta_policy
        names = [n for n, _ in meta.named_parameters()]
        self.meta_optimizer.zero_grad()
        loss_sum = torch.zeros((), dtype=torch.float64, device=self.device)
        lens, rews, query_stats = [], [], {}
        grads_sum = [torch.zeros_like(p) for p in meta.parameters()]

        if my_seeds:
            B = len(my_seeds)
            env = self._task_env(my_seeds)
            graphed = self.use_cuda_graph and self._lean(env, meta)
            # inner loop: support rollout under the shared meta weights, one SGD step per task
            support = self.collect_trajectory(env, meta, steps=k_support)

            def adapt():
                fast0 = _stack(meta, B)
                s_loss, _ = self._loss_terms(support, meta, fast0, want_stats=False)
                return self._inner_step(fast0, s_loss, names, self.lr_inner)
            # (the loss / gradient passes are launch-bound too -- a few hundred small kernels plus vmap's host-side
            # dispatch -- so each is captured once per (env, k) and replayed: the rollouts' buffers they read and the
            # per-task weights they write are the graphs' own static tensors)
            fast = self._phase(("adapt", id(env), k_support), adapt) if graphed else adapt()
            # outer loop: query rollout under each task's adapted weights, first-order gradient
            query = self.collect_trajectory(env, meta, steps=k_query, params=fast, params_static=graphed)

            def meta_grad():
                leaves = {n: fast[n].detach().requires_grad_(True) for n in names}
                q_loss, stat5 = self._loss_terms(query, meta, leaves)
                gq = torch.autograd.grad(q_loss.sum(), [leaves[n] for n in names])
                with torch.no_grad():  # keep `fast_policy` = the last task's adapted weights, as the reference leaves it
                    for n, p in self.fast_policy.named_parameters():
                        p.copy_(fast[n][-1])
                    sums = [gi.sum(0) for gi in gq]
                    return sums, q_loss.detach().double().sum(), torch.cat([stat5, q_loss.detach().mean().reshape(1)])
            grads_sum, loss_sum, stat6 = (self._phase(("meta_grad", id(env), k_query), meta_grad) if graphed
                                          else meta_grad())
            lens, rews = query["ep_lens"], query["ep_rews"]   # first host read-back of the iteration: everything is queued
            s6 = stat6.tolist()
            query_stats = {"pi_loss": s6[0], "v_loss": s6[1], "entropy": s6[2], "kl": s6[3], "clipfrac": s6[4],
                           "loss": s6[5]}

        # the one collective: SUM of the accumulated meta-gradient (and of the logged loss), then / global #tasks
        flat = torch.cat([gi.reshape(-1) for gi in grads_sum] + [loss_sum.reshape(1).to(grads_sum[0].dtype)])
        if parallel.world_size() > 1:
            torch.distributed.all_reduce(flat, op=torch.distributed.ReduceOp.SUM)
        flat = flat / n_global
        off = 0
        for p in meta.parameters():
            p.grad = flat[off: off + p.numel()].view_as(p).clone()
            off += p.numel()
        avg_loss = float(flat[off].item())
        torch.nn.utils.clip_grad_norm_(meta.parameters(), max_norm=0.5)
        self.meta_optimizer.step()

        if len(rews) > 0:
            avg_rew, avg_steps = float(np.mean(rews)), float(np.mean(lens))
        else:
            avg_rew, avg_steps = 0.0, float(k_query)
        return avg_loss, avg_rew, avg_steps, query_stats


class _Trajectory(dict):
    """The trajectory dict of `collect_trajectory`.  A fused rollout stores the 147-byte symbolic observations
    (`obs_symbolic`, `[steps, B, 7, 7, 3]`); the reference's `obs` entry -- the frames the policy saw,
    `[steps, B, 56, 56, 3]` uint8 -- is rendered from them the first time it is asked for."""
    env = None
    stats = None

    def __missing__(self, key):
        if key in ("ep_lens", "ep_rews") and self.stats is not None:
            ep_len, ep_ret = self.stats
            ended = ep_len > 0
            self["ep_lens"], self["ep_rews"] = ep_len[ended].tolist(), ep_ret[ended].tolist()
            return self[key]
        if key == "obs" and "obs_symbolic" in self and self.env is not None:
            sym = self["obs_symbolic"]
            frames = self.env.render(sym.reshape(-1, 7, 7, 3)).view(sym.shape[:2] + (56, 56, 3))
            self["obs"] = frames
            return frames
        raise KeyError(key)


def _stack(policy, B):
    """The module's parameters repeated B times along a new leading axis: one independent copy per task."""
    return {n: p.detach().unsqueeze(0).repeat((B,) + (1,) * p.dim()).requires_grad_(True)
            for n, p in policy.named_parameters()}


def _logits_value(policy, params, obs, **kw):
    """`policy._logits_value(obs)` evaluated with `params` substituted for the module's own weights."""
    return functional_call(policy, params, (obs,), kw)


def _blocked_kernel(conv1_weight):
    """One task's first-layer kernel re-indexed for space-to-depth input (CNNFeatureExtractor.blocked_weight)."""
    from .actor_critic import _space_to_depth4_weight
    return _space_to_depth4_weight(conv1_weight) * (1.0 / 255.0)


def _clip_coef(grads, max_norm):
    """Per-task `clip_grad_norm_` coefficient for stacked gradients `[B, ...]` (src/fomaml.py:181)."""
    sq = sum(g.reshape(g.shape[0], -1).pow(2).sum(1) for g in grads)
    return torch.clamp(max_norm / (sq.sqrt() + 1e-6), max=1.0)
