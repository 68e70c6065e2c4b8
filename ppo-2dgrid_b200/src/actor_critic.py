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
