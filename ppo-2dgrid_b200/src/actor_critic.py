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
# Inference-only evaluation for rollouts.  The parameters do not change while a rollout is collected, so the two
# Nature-CNN trunks can be packed once per rollout into ONE network of fused layers (the same arithmetic per output
# element: every sum runs over the same terms; cuDNN / cuBLAS may order them differently, which is immaterial for a
# path whose actions are sampled):
#   conv1    actor and critic kernels concatenated over the output channels        [64, 48, 2, 2]   (blocked input)
#   conv2/3  one grouped convolution, groups = 2 (actor channels | critic channels) [128, 32, 4, 4], [128, 64, 3, 3]
#   hidden   one batched matmul over the two trunks                                  [2, 576, 512]
#   heads    one batched matmul; the critic's single column is padded to the actor's width
# bias + ReLU ride in the convolution (torch.cudnn_convolution_relu) where cuDNN offers the fused engine.
# 8 kernels per policy evaluation instead of ~35, which is what a 32-task FOMAML rollout (launch-latency-bound) feels.
# With `params` (stacked per-task weights [B, ...], one frame per task) the same network runs as grouped convolutions
# with groups = B / 2B and batched matmuls over 2B (task, trunk) pairs.
class RolloutPolicy:
    """`logits, value = rp(x)` for `x` = float32 blocked frames `[N, 14, 14, 48]` holding pixel values 0..255
    (BatchedMerlinEnv.render(..., blocked=True, dtype=torch.float32)).  `logits` `[N, A]` and `value` `[N]` are
    strided views of one output tensor: hand them to `BatchedMerlinEnv.make_policy_io` as they are.
    Built from a CNNActorCritic (shared weights) or from stacked per-task parameters (`params`, leading axis B: task b
    is evaluated on frame b with its own weights).  Call `refresh()` after the weights changed."""

    _fused_conv_ok = None  # probed once per process: does cuDNN run the fused conv+bias+ReLU for these shapes?

    def __init__(self, ac, params=None, use_fused_conv=True):
        self.ac, self.params = ac, params
        self.per_task = params is not None
        self.act_dim = ac.actor[-1].out_features
        self.use_fused_conv = use_fused_conv
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
            w1 = torch.cat([_space_to_depth4_weight(g(t, 0, "weight")) for t in ("actor", "critic")]) * (1.0 / 255.0)
            self.w1 = w1.contiguous(memory_format=torch.channels_last)
            self.b1 = torch.cat([g(t, 0, "bias") for t in ("actor", "critic")]).contiguous()
            self.w2 = torch.cat([g(t, 2, "weight") for t in ("actor", "critic")]).contiguous(memory_format=torch.channels_last)
            self.b2 = torch.cat([g(t, 2, "bias") for t in ("actor", "critic")]).contiguous()
            self.w3 = torch.cat([g(t, 4, "weight") for t in ("actor", "critic")]).contiguous(memory_format=torch.channels_last)
            self.b3 = torch.cat([g(t, 4, "bias") for t in ("actor", "critic")]).contiguous()
            # hidden layer: [2, 576, 512] acting on features in the (h, w, c) order the channels-last conv output has
            def hid(head):
                w = h(head, 0, "weight")  # [512, 64*3*3] over (c, h, w)
                return w.view(-1, 64, 3, 3).permute(2, 3, 1, 0).reshape(576, -1)
            self.wh = torch.stack([hid("actor"), hid("critic")]).contiguous()
            self.bh = torch.stack([h("actor", 0, "bias"), h("critic", 0, "bias")]).unsqueeze(1).contiguous()  # [2, 1, 512]
            wo = torch.zeros((2, self.wh.shape[2], A), dtype=w1.dtype, device=w1.device)
            wo[0] = h("actor", 2, "weight").t()
            wo[1, :, :1] = h("critic", 2, "weight").t()
            bo = torch.zeros((2, 1, A), dtype=w1.dtype, device=w1.device)
            bo[0, 0] = h("actor", 2, "bias")
            bo[1, 0, :1] = h("critic", 2, "bias")
            self.wo, self.bo = wo, bo
            self.groups = (1, 2, 2)
        else:
            B = g("actor", 0, "weight").shape[0]
            self.B = B
            def conv(i, blocked=False):
                wa, wc = g("actor", i, "weight"), g("critic", i, "weight")  # [B, O, C, k, k]
                if blocked:
                    o, c = wa.shape[1], wa.shape[2]
                    f = lambda w: (w.reshape(B, o, c, 2, 4, 2, 4).permute(0, 1, 2, 4, 6, 3, 5)  # noqa: E731
                                   .reshape(B, o, c * 16, 2, 2) * (1.0 / 255.0))
                    wa, wc = f(wa), f(wc)
                w = torch.stack([wa, wc], 1)  # [B, 2, O, C, k, k] -> groups ordered (task, trunk)
                bias = torch.stack([g("actor", i, "bias"), g("critic", i, "bias")], 1)  # [B, 2, O]
                return w.reshape((-1,) + tuple(w.shape[3:])).contiguous(), bias.reshape(-1).contiguous()
            self.w1, self.b1 = conv(0, blocked=True)   # [B*64, 48, 2, 2], groups B (both trunks read the task's frame)
            self.w2, self.b2 = conv(2)                 # [B*128, 32, 4, 4], groups 2B
            self.w3, self.b3 = conv(4)                 # [B*128, 64, 3, 3], groups 2B
            self.wh = torch.stack([h("actor", 0, "weight"), h("critic", 0, "weight")], 1).reshape(2 * B, -1, 576).transpose(1, 2)
            self.bh = torch.stack([h("actor", 0, "bias"), h("critic", 0, "bias")], 1).reshape(2 * B, 1, -1).contiguous()
            hd = self.bh.shape[2]
            wo = torch.zeros((B, 2, hd, A), dtype=self.w1.dtype, device=self.w1.device)
            wo[:, 0] = h("actor", 2, "weight").transpose(1, 2)
            wo[:, 1, :, :1] = h("critic", 2, "weight").transpose(1, 2)
            bo = torch.zeros((B, 2, 1, A), dtype=self.w1.dtype, device=self.w1.device)
            bo[:, 0, 0] = h("actor", 2, "bias")
            bo[:, 1, 0, :1] = h("critic", 2, "bias")
            self.wo, self.bo = wo.reshape(2 * B, hd, A), bo.reshape(2 * B, 1, A)
            self.groups = (B, 2 * B, 2 * B)

    def _conv(self, x, w, b, stride, groups):
        if self.use_fused_conv and RolloutPolicy._fused_conv_ok is not False:
            try:
                y = torch.cudnn_convolution_relu(x, w, b, (stride, stride), (0, 0), (1, 1), groups)
                RolloutPolicy._fused_conv_ok = True
                return y
            except RuntimeError:
                if RolloutPolicy._fused_conv_ok:  # it worked before: a real error, not a missing engine
                    raise
                RolloutPolicy._fused_conv_ok = False
        return torch.relu_(torch.nn.functional.conv2d(x, w, b, stride=stride, groups=groups))

    @torch.no_grad()
    def __call__(self, x, out=None):
        """`out`: optional static `[2, N, A]` (shared weights) / `[2B, 1, A]` (per-task) tensor receiving the heads."""
        A = self.act_dim
        if not self.per_task:
            n = x.shape[0]
            y = x.permute(0, 3, 1, 2)  # [N, 48, 14, 14], channels-last memory as rendered
            y = self._conv(y, self.w1, self.b1, 1, 1)
            y = self._conv(y, self.w2, self.b2, 2, 2)
            y = self._conv(y, self.w3, self.b3, 1, 2)       # [N, 128, 3, 3]
            f = y.permute(0, 2, 3, 1).reshape(n, 9, 2, 64).permute(2, 0, 1, 3).reshape(2, n, 576)  # (trunk, n, (h, w, c))
            hid = torch.relu_(torch.baddbmm(self.bh, f, self.wh))          # [2, N, 512]
            out = torch.baddbmm(self.bo, hid, self.wo, out=out)            # [2, N, A]
            return out[0], out[1, :, 0]
        B = self.B
        y = x.permute(0, 3, 1, 2).reshape(1, B * 48, 14, 14)  # one copy: (task, channel) become grouped channels
        y = self._conv(y, self.w1, self.b1, 1, self.groups[0])
        y = self._conv(y, self.w2, self.b2, 2, self.groups[1])
        y = self._conv(y, self.w3, self.b3, 1, self.groups[2])   # [1, B*128, 3, 3]
        f = y.reshape(2 * B, 1, 576)                                # (task, trunk) x (c, h, w)
        hid = torch.relu_(torch.baddbmm(self.bh, f, self.wh))      # [2B, 1, 512]
        out = torch.baddbmm(self.bo, hid, self.wo, out=out).view(B, 2, A)   # [B, 2, A]
        return out[:, 0], out[:, 1, 0]
