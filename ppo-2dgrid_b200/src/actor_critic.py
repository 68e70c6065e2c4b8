"""Actor-critic networks (PyTorch; out of the CUDA hot-path scope, kept so the drop-in is runnable).

Architecture, initialisation and state_dict keys follow the reference (src/actor_critic.py:5-99) so that
checkpoints are interchangeable: two separate Nature-CNN trunks (conv 8/4 -> 4/2 -> 3/1, ReLU) + 512-wide
heads for images `[N, H, W, 3]`, a 64-64 tanh MLP pair for flat observations.  Unlike the reference the
image path accepts the device-resident uint8 frames of the batched env directly (cast + /255 on the fly).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .utils.utils_rl import layer_init


def _conv_stack(channels):
    return nn.Sequential(
        layer_init(nn.Conv2d(channels, 32, kernel_size=8, stride=4)), nn.ReLU(),
        layer_init(nn.Conv2d(32, 64, kernel_size=4, stride=2)), nn.ReLU(),
        layer_init(nn.Conv2d(64, 64, kernel_size=3, stride=1)), nn.ReLU(),
        nn.Flatten(),
    )


def space_to_depth4(frames):
    """`[N, H, W, C]` frames (uint8 from the env, or float) -> float `[N, 16*C, H/4, W/4]` in channels-last memory:
    every 4x4 pixel block becomes one position with channel index `c*16 + dy*4 + dx`."""
    n, h, w, c = frames.shape
    x = frames.reshape(n, h // 4, 4, w // 4, 4, c).permute(0, 1, 3, 5, 2, 4).reshape(n, h // 4, w // 4, c * 16)
    return x.permute(0, 3, 1, 2).float()


def _space_to_depth4_weight(weight):
    """The `[O, C, 8, 8]` stride-4 kernel as the equivalent `[O, 16*C, 2, 2]` stride-1 kernel over space_to_depth4."""
    o, c, _, _ = weight.shape
    return weight.reshape(o, c, 2, 4, 2, 4).permute(0, 1, 3, 5, 2, 4).reshape(o, c * 16, 2, 2)


class CNNFeatureExtractor(nn.Module):
    def __init__(self, channels, height, width):
        super().__init__()
        self.network = _conv_stack(channels)
        with torch.no_grad():
            self.output_dim = self.network(torch.zeros(1, channels, height, width)).shape[1]
        self.blockable = height % 4 == 0 and width % 4 == 0

    def forward(self, x):
        """x: `[N, C, H, W]` float pixel values in 0..255 (the reference's input convention)."""
        return self.network(x / 255.0)

    def blocked_weight(self):
        """conv1's `[32, 3, 8, 8]` stride-4 kernel re-indexed as the `[32, 48, 2, 2]` stride-1 kernel over
        `space_to_depth4` input, 1/255 folded in.  A rollout can form it once (`CNNActorCritic.blocked_weights`)."""
        return _space_to_depth4_weight(self.network[0].weight) * (1.0 / 255.0)

    def forward_blocked(self, xb, weight=None):
        """Same function on `space_to_depth4` input (float pixel values 0..255).  The first layer (8x8, stride 4, 3
        input channels) is evaluated as a 2x2 stride-1 convolution over 48 channels with the SAME parameters
        (re-indexed on the fly, 1/255 folded into them): identical sums in a different order, and a shape cuDNN runs
        several times faster than C = 3."""
        w = self.blocked_weight() if weight is None else weight
        h = torch.nn.functional.conv2d(xb, w, self.network[0].bias)
        return self.network[1:](h)


def _head(in_dim, hidden, out_dim, out_std, act):
    return nn.Sequential(layer_init(nn.Linear(in_dim, hidden)), act(), layer_init(nn.Linear(hidden, out_dim), std=out_std))


class _ActorCriticBase(nn.Module):
    """Shared act/evaluate on top of `_logits_value(obs)`.  Sampling, log-probabilities and entropy are the
    categorical-distribution formulas written out on tensors (no distribution object, no argument validation),
    so that a whole rollout can be captured in a CUDA graph without a host synchronisation."""

    def forward(self, obs, **kw):
        """(logits `[B, A]`, value `[B]`) -- the functional entry point used for stacked per-task weights."""
        return self._logits_value(obs, **kw)

    def act(self, obs, deterministic=False, **kw):
        logits, value = self._logits_value(obs, **kw)
        logp_all = torch.log_softmax(logits, dim=-1)
        if deterministic:
            action = torch.argmax(logits, dim=1)
        else:
            action = torch.multinomial(logp_all.exp(), 1).squeeze(-1)
        return action, logp_all.gather(-1, action.unsqueeze(-1)).squeeze(-1), value

    def evaluate(self, obs, actions):
        logits, value = self._logits_value(obs)
        logp_all = torch.log_softmax(logits, dim=-1)
        entropy = -(logp_all.exp() * logp_all).sum(-1)
        return logp_all.gather(-1, actions.long().unsqueeze(-1)).squeeze(-1), entropy, value


class CNNActorCritic(_ActorCriticBase):
    def __init__(self, obs_shape, act_dim, hidden_dim=512):
        super().__init__()
        h, w, c = obs_shape
        self.actor_extractor = CNNFeatureExtractor(c, h, w)
        self.critic_extractor = CNNFeatureExtractor(c, h, w)
        self.actor = _head(self.actor_extractor.output_dim, hidden_dim, act_dim, 0.01, nn.ReLU)
        self.critic = _head(self.critic_extractor.output_dim, hidden_dim, 1, 1.0, nn.ReLU)

        self.blocked_first_layer = self.actor_extractor.blockable  # False: evaluate conv1 literally (8x8, stride 4)

    def _format_obs(self, x):
        if x.ndim == 4 and x.shape[-1] == 3:  # NHWC (uint8 frames from the env, or float copies) -> NCHW float
            return x.permute(0, 3, 1, 2).float()
        return x.float()

    def blocked_weights(self):
        """The two trunks' re-indexed first-layer kernels, for `act(..., blocked=...)`: a rollout whose parameters do
        not change between steps forms them once instead of at every step (four small kernels per step)."""
        return self.actor_extractor.blocked_weight(), self.critic_extractor.blocked_weight()

    def _logits_value(self, obs, blocked=None):
        wa, wc = blocked if blocked is not None else (None, None)
        if obs.ndim == 4 and obs.shape[-1] == 48:
            # frames already in the blocked layout [N, H/4, W/4, 48] (BatchedMerlinEnv.render(..., blocked=True)), pixel
            # values 0..255 as uint8 or -- written by the render kernel itself, nothing to cast or copy here -- float32
            xb = obs.permute(0, 3, 1, 2)
            if xb.dtype != torch.float32:
                xb = xb.float()
            fa, fc = self.actor_extractor.forward_blocked(xb, wa), self.critic_extractor.forward_blocked(xb, wc)
        elif self.blocked_first_layer and obs.ndim == 4 and obs.shape[-1] == 3:
            xb = space_to_depth4(obs)  # shared by both trunks
            fa, fc = self.actor_extractor.forward_blocked(xb, wa), self.critic_extractor.forward_blocked(xb, wc)
        else:
            obs = self._format_obs(obs)
            fa, fc = self.actor_extractor(obs), self.critic_extractor(obs)
        return self.actor(fa), self.critic(fc).squeeze(-1)


class MLPActorCritic(_ActorCriticBase):
    def __init__(self, obs_dim, act_dim, hidden_dim=64):
        super().__init__()

        def tower(out_dim, out_std):
            return nn.Sequential(layer_init(nn.Linear(obs_dim, hidden_dim)), nn.Tanh(),
                                 layer_init(nn.Linear(hidden_dim, hidden_dim)), nn.Tanh(),
                                 layer_init(nn.Linear(hidden_dim, out_dim), std=out_std))

        self.actor = tower(act_dim, 0.01)
        self.critic = tower(1, 1.0)

    def _logits_value(self, obs):
        obs = obs.float()
        return self.actor(obs), self.critic(obs).squeeze(-1)


# ---------------------------------------------------------------------------------------------------------------
# Inference-only evaluation for rollouts.  The parameters do not change while a rollout is collected, so the network
# can be re-packed once per rollout into the form that costs the fewest launches -- a 32-task FOMAML rollout is 512
# sequential policy evaluations of a few microseconds of arithmetic each, i.e. launch-latency-bound.  Same arithmetic
# per output element (every sum runs over the same terms; cuDNN / cuBLAS may order them differently, immaterial for
# a path whose actions are sampled).
#   shared weights   per trunk three cuDNN convolutions with bias + ReLU fused in (torch.cudnn_convolution_relu,
#                    channels-last, first layer on the blocked input with 1/255 folded in), the hidden layer as one
#                    GEMM with bias + ReLU in its epilogue (torch._addmm_activation), the head as one GEMM writing
#                    straight into the static `[2, N, A]` tensor the fused env transition reads: 10 launches.
#                    (Grouped convolutions over both trunks were tried and are slower: cuDNN falls back to direct /
#                    SGEMM engines plus layout-conversion kernels for them.)
#   per-task weights (stacked `[B, ...]`, one frame per task: FOMAML query rollouts, few-shot evaluation) every layer is
#                    ONE batched matmul over tasks (x trunks): the convolutions as im2col + bmm -- the patches are an
#                    `unfold` view copied once into a static buffer whose extra column of ones carries the bias -- so
#                    no grouped convolution with 32 / 64 groups is involved: 14 launches instead of ~70 under vmap.
class RolloutPolicy:
    """`logits, value = rp(x)` for `x` = float32 blocked frames `[N, 14, 14, 48]` holding pixel values 0..255
    (BatchedMerlinEnv.render(..., blocked=True, dtype=torch.float32)).  `logits` `[N, A]` and `value` `[N]` are
    strided views of one output tensor: hand them to `BatchedMerlinEnv.make_policy_io` as they are.
    Built from a CNNActorCritic (shared weights) or from stacked per-task parameters (`params`, leading axis B: task b
    is evaluated on frame b with its own weights).  Call `refresh()` after the weights changed."""

    _fused_conv_ok = None  # probed once per process: does cuDNN run the fused conv+bias+ReLU for these shapes?

    def __init__(self, ac, params=None, use_fused_conv=True, actor_only=False):
        self.ac, self.params = ac, params
        self.per_task = params is not None
        # actor_only (per-task weights): greedy evaluation needs no value -- the critic trunk's half of every batched
        # matmul (and of the per-task weight traffic, ~150 MB per step at 100 tasks) is skipped; `value` is then None
        self.trunks = ("actor",) if (actor_only and self.per_task) else ("actor", "critic")
        self.act_dim = ac.actor[-1].out_features
        self.use_fused_conv = use_fused_conv
        self._patch = None   # per-task path: static im2col buffers (bias column pre-filled)
        self._hid = None     # shared path: static [2, N, 512] hidden activations
        self._side = None    # shared path: side stream the critic trunk runs on for small batches
        self.fork_below = 2048
        self.refresh()

    def _p(self, name):
        if self.params is not None:
            return self.params[name]
        return dict(self.ac.named_parameters())[name]

    @torch.no_grad()
    def refresh(self):
        A = self.act_dim
        g = lambda trunk, i, kind: self._p(f"{trunk}_extractor.network.{i}.{kind}")  # noqa: E731
        h = lambda head, i, kind: self._p(f"{head}.{i}.{kind}")  # noqa: E731
        if not self.per_task:
            cl = torch.channels_last
            self.conv = {}
            for t in ("actor", "critic"):
                self.conv[t] = [((_space_to_depth4_weight(g(t, 0, "weight")) * (1.0 / 255.0)).contiguous(memory_format=cl),
                                 g(t, 0, "bias").detach(), 1),
                                (g(t, 2, "weight").detach().contiguous(memory_format=cl), g(t, 2, "bias").detach(), 2),
                                (g(t, 4, "weight").detach().contiguous(memory_format=cl), g(t, 4, "bias").detach(), 1)]
            # hidden layers act on features in the (h, w, c) order the channels-last conv output has: [576, 512 + 8] --
            # unit 512 is a constant one (zero weights, bias 1: ReLU keeps it) that carries the heads' biases, 513.. are padding
            hd = h("actor", 0, "weight").shape[0]
            dev, dt = h("actor", 0, "weight").device, h("actor", 0, "weight").dtype
            self.wh, self.bh = {}, {}
            for t in ("actor", "critic"):
                w = torch.zeros((576, hd + 8), dtype=dt, device=dev)
                w[:, :hd] = h(t, 0, "weight").view(-1, 64, 3, 3).permute(2, 3, 1, 0).reshape(576, hd)
                bias = torch.zeros(hd + 8, dtype=dt, device=dev)
                bias[:hd] = h(t, 0, "bias")
                bias[hd:hd + 1].fill_(1.0)   # (an integer-indexed `bias[hd] = 1.0` is a host-to-device copy: illegal under capture)
                self.wh[t], self.bh[t] = w, bias
            # both heads over the trunks at once, bias in row 512; the critic's single column is padded to the actor's width
            self.wo = torch.zeros((2, hd + 8, A), dtype=dt, device=dev)          # for a batched matmul
            self.wo[0, :hd] = h("actor", 2, "weight").t()
            self.wo[0, hd] = h("actor", 2, "bias")
            self.wo[1, :hd, :1] = h("critic", 2, "weight").t()
            self.wo[1, hd, :1] = h("critic", 2, "bias")
            self.wo_t = self.wo.transpose(1, 2).contiguous().unsqueeze(1)       # [2, 1, A, 520] for multiply-and-reduce
            return
        B = g("actor", 0, "weight").shape[0]
        self.B = B
        tr, T = self.trunks, len(self.trunks)
        dev, dt = g("actor", 0, "weight").device, g("actor", 0, "weight").dtype

        def aug(w, bias, k_pad):
            """[G, K, O] weights + [G, O] bias -> [G, k_pad, O]: the bias as row K (the patches' column of ones), zeros after."""
            G, K, O = w.shape
            out = torch.zeros((G, k_pad, O), dtype=dt, device=dev)
            out[:, :K] = w
            out[:, K] = bias
            return out

        # conv1 on the blocked input: [B, 192, 32 T] -- the trunks side by side on the output axis, 1/255 folded in
        def blocked(w):  # [B, 32, 3, 8, 8] -> [B, 32, 192] over (c*16 + dy*4 + dx, by, bx) = the blocked channel, 2x2 taps
            return w.reshape(B, 32, 3, 2, 4, 2, 4).permute(0, 1, 2, 4, 6, 3, 5).reshape(B, 32, 192) * (1.0 / 255.0)
        w1 = torch.cat([blocked(g(t, 0, "weight")) for t in tr], 1).transpose(1, 2)
        self.w1 = aug(w1, torch.cat([g(t, 0, "bias") for t in tr], 1), 196)
        # conv2 / conv3 per (task, trunk): [T B, C*k*k, 64]
        def per_trunk(i):
            w = torch.stack([g(t, i, "weight") for t in tr], 1)   # [B, T, 64, C, k, k]
            bias = torch.stack([g(t, i, "bias") for t in tr], 1)    # [B, T, 64]
            w = w.reshape(T * B, 64, -1).transpose(1, 2)
            return aug(w, bias.reshape(T * B, 64), w.shape[1] + 4)
        self.w2, self.w3 = per_trunk(2), per_trunk(4)     # [T B, 516, 64], [T B, 580, 64]
        # hidden layer on features in (h, w, c) order: [T B, 576, 512 + 4] -- column 512 is a constant-one unit (zero
        # weights, bias 1) that carries the heads' biases, columns 513.. are zero padding
        wh = torch.stack([h(t, 0, "weight") for t in tr], 1)   # [B, T, 512, 64*3*3] over (c, h, w)
        hd = wh.shape[2]
        self.wh = torch.zeros((T * B, 576, hd + 4), dtype=dt, device=dev)
        self.wh[:, :, :hd] = wh.reshape(T * B, hd, 64, 9).permute(0, 3, 2, 1).reshape(T * B, 576, hd)
        self.bh = torch.zeros((T * B, 1, hd + 4), dtype=dt, device=dev)
        self.bh[:, 0, :hd] = torch.stack([h(t, 0, "bias") for t in tr], 1).reshape(T * B, hd)
        self.bh[:, 0, hd].fill_(1.0)
        # heads, transposed for a multiply-and-reduce (a [T B, 1, 512] x [T B, 512, 3] bmm costs 17 us as a batched GEMV):
        # [T B, A, 512 + 4] with the bias in column 512; the critic's single row is padded to the actor's width
        wo = torch.zeros((B, T, A, hd + 4), dtype=dt, device=dev)
        wo[:, 0, :, :hd] = h("actor", 2, "weight")
        wo[:, 0, :, hd] = h("actor", 2, "bias")
        if T == 2:
            wo[:, 1, :1, :hd] = h("critic", 2, "weight")
            wo[:, 1, :1, hd] = h("critic", 2, "bias")
        self.wo = wo.reshape(T * B, A, hd + 4)
        if self._patch is None:
            def ones_col(G, L, K):
                buf = torch.zeros((G, L, K + 4), dtype=dt, device=dev)
                buf[:, :, K].fill_(1.0)
                return buf
            self._patch = (ones_col(B, 169, 192), ones_col(T * B, 25, 512), ones_col(T * B, 9, 576))

    def _conv(self, x, w, b, stride):
        if self.use_fused_conv and RolloutPolicy._fused_conv_ok is not False:
            try:
                y = torch.cudnn_convolution_relu(x, w, b, (stride, stride), (0, 0), (1, 1), 1)
                RolloutPolicy._fused_conv_ok = True
                return y
            except RuntimeError:
                if RolloutPolicy._fused_conv_ok:  # it worked before: a real error, not a missing engine
                    raise
                RolloutPolicy._fused_conv_ok = False
        return torch.relu_(torch.nn.functional.conv2d(x, w, b, stride=stride))

    @torch.no_grad()
    def __call__(self, x, out=None):
        """`out`: optional static `[2, N, A]` (shared weights) / `[2B, 1, A]` (per-task) tensor receiving the heads."""
        A = self.act_dim
        if not self.per_task:
            n = x.shape[0]
            if out is None:
                out = torch.empty((2, n, A), dtype=x.dtype, device=x.device)
            xc = x.permute(0, 3, 1, 2)  # [N, 48, 14, 14], channels-last memory as rendered
            if self._hid is None or self._hid.shape[1] != n:
                self._hid = torch.empty((2, n, self.wh["actor"].shape[1]), dtype=x.dtype, device=x.device)
            def trunk(i, t):
                y = xc
                for w, b, stride in self.conv[t]:
                    y = self._conv(y, w, b, stride)                       # ... [N, 64, 3, 3]
                f = y.permute(0, 2, 3, 1).reshape(n, 576)                  # (h, w, c): a view of the channels-last output
                torch._addmm_activation(self.bh[t], f, self.wh[t], out=self._hid[i])   # bias + ReLU in the GEMM epilogue

            if x.is_cuda and n <= self.fork_below:
                # small batch = latency-bound: the two trunks are independent chains of four tiny kernels; the critic's
                # runs on a side stream beside the actor's (fork / join; captured as parallel branches of a CUDA graph)
                cur = torch.cuda.current_stream(x.device)
                if self._side is None:
                    self._side = torch.cuda.Stream(x.device)
                self._side.wait_stream(cur)
                with torch.cuda.stream(self._side):
                    trunk(1, "critic")
                trunk(0, "actor")
                cur.wait_stream(self._side)
            else:
                trunk(0, "actor")
                trunk(1, "critic")
            if n <= 256:   # a [N, 520] x [520, 3] matmul costs 11 us as a GEMM at these sizes; multiply-and-reduce: 4 us
                torch.sum(self._hid.unsqueeze(2) * self.wo_t, dim=-1, out=out)
            else:
                torch.bmm(self._hid, self.wo, out=out)                      # [2, N, A]
            return out[0], out[1, :, 0]
        B, T = self.B, len(self.trunks)
        p1, p2, p3 = self._patch
        # the three convolutions run in the precision PyTorch runs convolutions in (TF32 tensor cores unless
        # torch.backends.cudnn.allow_tf32 was switched off) -- what cuDNN does for the same layers in the loss and in the
        # shared-weight path; the linear layers below stay in fp32 like nn.Linear
        mm = torch.backends.cuda.matmul
        keep, mm.allow_tf32 = mm.allow_tf32, bool(torch.backends.cudnn.allow_tf32) or mm.allow_tf32
        try:
            # conv1: 2x2 taps over the 14x14x48 blocked frame -> 13x13 positions x 192
            # (the patch buffers are viewed in the unfolded shape, so each im2col is ONE strided copy, no intermediate)
            p1[:, :, :192].unflatten(2, (48, 2, 2)).unflatten(1, (13, 13)).copy_(x.unfold(1, 2, 1).unfold(2, 2, 1))
            h1 = torch.relu_(torch.bmm(p1, self.w1))                       # [B, 169, 32 T] = [B, 13, 13, (trunk, 32)]
            # conv2: 4x4 stride 2 over 13x13x32 per trunk -> 5x5 positions x 512
            v = h1.view(B, 13, 13, T, 32).unfold(1, 4, 2).unfold(2, 4, 2)  # [B, 5, 5, T, 32, 4, 4]
            p2[:, :, :512].unflatten(2, (32, 4, 4)).unflatten(1, (5, 5)).unflatten(0, (B, T)).copy_(v.permute(0, 3, 1, 2, 4, 5, 6))
            h2 = torch.relu_(torch.bmm(p2, self.w2))                       # [T B, 25, 64] = [T B, 5, 5, 64]
            # conv3: 3x3 over 5x5x64 -> 3x3 positions x 576
            v = h2.view(T * B, 5, 5, 64).unfold(1, 3, 1).unfold(2, 3, 1)   # [T B, 3, 3, 64, 3, 3]
            p3[:, :, :576].unflatten(2, (64, 3, 3)).unflatten(1, (3, 3)).copy_(v)
            h3 = torch.relu_(torch.bmm(p3, self.w3))                       # [T B, 9, 64]
        finally:
            mm.allow_tf32 = keep
        hid = torch.relu_(torch.baddbmm(self.bh, h3.view(T * B, 1, 576), self.wh))    # [T B, 1, 512 + 4]
        if out is None:
            out = torch.empty((T * B, 1, A), dtype=x.dtype, device=x.device)
        torch.sum(hid * self.wo, dim=-1, out=out.view(T * B, A))           # heads (+ bias through the constant-one unit)
        out = out.view(B, T, A)
        return out[:, 0], (out[:, 1, 0] if T == 2 else None)
