"""ORACLE (test infrastructure -- imported only by tests/, __graft_entry__.smoke() and bench.py's CPU legs).

CPU restatement of the in-kernel action sampler of merlin_env_policy_step (include/merlin_b200.h).  The reference
samples with torch.distributions.Categorical(logits).sample() and stores .log_prob(a) (src/actor_critic.py act();
src/ppo.py:70-86; src/fomaml.py:65-84); torch's generator stream cannot be reproduced by a fused kernel, so the draw is
SPECIFIED by the product and restated here independently (pure-Python integer arithmetic for Philox, numpy float32 for
the inverse CDF):

    u      = (Philox4x32-10(counter = (env, draw, 0, 0), key = (seed_lo, seed_hi))[0] >> 8) * 2**-24
    action = first a with cumsum(exp(l - max(l)))[a] > u * sum(exp(l - max(l)))          (float32, left to right)
    logp   = (l[action] - max(l)) - log(sum)

Philox4x32-10: Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3", SC'11 (Random123).  Pinned by
the Random123 known-answer vectors in tests/test_host_logic.py.  exp/log differ in the last ulp between libm and the
GPU's expf/logf: tests compare log-probabilities with a tolerance and skip actions whose draw lies within that
tolerance of a CDF boundary.
"""
from __future__ import annotations

import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(counter, key):
    """counter: 4 ints, key: 2 ints (all uint32) -> 4 uint32 words."""
    c0, c1, c2, c3 = [int(c) & MASK for c in counter]
    k0, k1 = [int(k) & MASK for k in key]
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0, c1, c2, c3


def uniform(seed, env, draw):
    """The float32 uniform in [0, 1) env `env` uses for its draw number `draw`."""
    x = philox4x32_10((env, draw, 0, 0), (seed & MASK, (seed >> 32) & MASK))[0]
    return np.float32(x >> 8) * np.float32(1.0 / 16777216.0)


def sample(logits, u, greedy=False):
    """logits: float32[A]; returns (action, logp float32, margin) -- margin = distance of u*sum to the nearest CDF
    boundary relative to sum (a draw with a tiny margin may legitimately differ by one action between exp
    implementations)."""
    l = np.asarray(logits, dtype=np.float32)
    m = l.max()
    ex = np.exp((l - m).astype(np.float32)).astype(np.float32)
    s = np.float32(0.0)
    cum = []
    for e in ex:  # float32, left to right, as the kernel accumulates
        s = np.float32(s + e)
        cum.append(s)
    if greedy:
        a = int(np.argmax(l))  # first maximum
        margin = 1.0
    else:
        target = np.float32(np.float32(u) * s)
        a = len(l) - 1
        for i, c in enumerate(cum):
            if c > target:
                a = i
                break
        margin = float(min(abs(float(c) - float(target)) for c in cum[:-1]) / float(s)) if len(cum) > 1 else 1.0
    logp = np.float32(np.float32(l[a] - m) - np.float32(np.log(s)))
    return a, logp, margin


def sample_batch(logits, seed, draws, greedy=False):
    """logits float32[N, A], draws int[N] -> (actions int64[N], logp float32[N], margin float64[N])."""
    logits = np.asarray(logits, dtype=np.float32)
    N = logits.shape[0]
    act = np.zeros(N, dtype=np.int64)
    lp = np.zeros(N, dtype=np.float32)
    mg = np.zeros(N, dtype=np.float64)
    for e in range(N):
        u = np.float32(0.0) if greedy else uniform(seed, e, int(draws[e]))
        act[e], lp[e], mg[e] = sample(logits[e], u, greedy)
    return act, lp, mg
