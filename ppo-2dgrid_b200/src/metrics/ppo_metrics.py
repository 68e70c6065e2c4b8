"""Scalar logging helpers with the reference's names (src/metrics/ppo_metrics.py:7-57)."""
from __future__ import annotations

_KEYS = ("pi_loss", "v_loss", "entropy", "kl", "clipfrac", "gradnorm")


def aggregate_ppo_update_metrics(total_pi, total_v, total_ent, total_kl, total_clip, total_gnorm, nbatches):
    totals = (total_pi, total_v, total_ent, total_kl, total_clip, total_gnorm)
    if nbatches == 0:
        return {k: 0.0 for k in _KEYS}
    return {k: t / nbatches for k, t in zip(_KEYS, totals)}


def compute_episode_stats(episode_returns, episode_lengths):
    if len(episode_returns) == 0:
        return {"episode_return_mean": 0.0, "episode_length_mean": 0.0}
    return {"episode_return_mean": sum(episode_returns) / len(episode_returns),
            "episode_length_mean": sum(episode_lengths) / len(episode_lengths)}
