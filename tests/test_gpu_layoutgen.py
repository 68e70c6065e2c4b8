"""GPU tier: on-device layout generation (merlin_env_generate_layouts).  It cannot be bit-exact with numpy's PCG64
stream, so it is validated by (1) the structural invariants of each `_gen_grid` routine, (2) statistics against the
host generators (which ARE pinned to the reference by fixtures), (3) determinism / stream addressing, and (4) driving
the oracle with the generated pool: rollouts over device-made layouts stay bit-exact."""
from __future__ import annotations

from collections import deque

import numpy as np
import pytest
import torch

from oracle import fast

pytestmark = pytest.mark.gpu

E, WALL, GOAL = 1, 2 | (5 << 4), 8 | (1 << 4)


def _env(N, size, difficulty, seed, L, **kw):
    from merlin_b200 import BatchedMerlinEnv
    return BatchedMerlinEnv(N, width=size, height=size, device="cuda:0", generate=(difficulty, seed, L), **kw)


def _reachable(g, agent, goal):
    H, W = g.shape
    seen = np.zeros_like(g, dtype=bool)
    q = deque([(agent[0], agent[1])])
    seen[agent[1], agent[0]] = True
    while q:
        x, y = q.popleft()
        if (x, y) == goal:
            return True
        for nx, ny in ((x, y + 1), (x + 1, y), (x, y - 1), (x - 1, y)):
            if 0 <= nx < W and 0 <= ny < H and not seen[ny, nx] and g[ny, nx] in (E, GOAL):
                seen[ny, nx] = True
                q.append((nx, ny))
    return False


@pytest.mark.parametrize("difficulty,size", [("easy", 16), ("medium", 16), ("mediumhard", 16), ("hard", 16),
                                             ("hardest", 16), ("mediumhard", 32), ("hard", 32), ("hardest", 24),
                                             ("hard", 8), ("mediumhard", 8)])
def test_structural_invariants(difficulty, size):
    L = 600
    env = _env(8, size, difficulty, 12345, L)
    cells, agent = env.layouts_numpy()
    assert cells.shape == (L, size * size) and agent.shape == (L, 3)
    W = H = size
    mid = W // 2
    n_fallback = 0
    for l in range(L):
        g = cells[l].reshape(H, W)
        ax, ay, ad = agent[l]
        assert set(np.unique(g)) <= {E, WALL, GOAL}
        assert (g[0] == WALL).all() and (g[-1] == WALL).all() and (g[:, 0] == WALL).all() and (g[:, -1] == WALL).all()
        assert int((g == GOAL).sum()) == 1
        gy, gx = map(int, np.argwhere(g == GOAL)[0])
        assert 1 <= ax < W - 1 and 1 <= ay < H - 1 and 0 <= ad <= 3
        # easy puts the goal at a fixed cell AFTER placing the agent (easy_env.py:27-36): the agent may start on it
        assert g[ay, ax] == E or (difficulty == "easy" and g[ay, ax] == GOAL)
        assert _reachable(g, (ax, ay), (gx, gy)), (difficulty, l)
        inner_walls = int((g[1:-1, 1:-1] == WALL).sum())
        if difficulty == "easy":
            assert (gx, gy) == (W - 5, H - 5) and inner_walls == 0
        elif difficulty == "medium":
            assert inner_walls == 0
        elif difficulty == "mediumhard":
            interior = (W - 2) * (H - 2)
            lo, hi = max(1, int(interior * 0.10)), max(1, int(interior * 0.20))
            if inner_walls == 0:
                n_fallback += 1
            else:
                assert lo <= inner_walls <= hi, inner_walls
        elif difficulty == "hard":
            col = g[1:-1, mid]
            gaps = int((col != WALL).sum())
            if gaps == H - 2:  # empty-room fallback
                n_fallback += 1
                continue
            assert (2 <= gaps <= 5) if W > 10 else gaps == 1
            assert ax < mid < gx
            extra = inner_walls - int((col == WALL).sum())
            assert (0 <= extra <= 12) if W > 10 else extra == 0
        elif difficulty == "hardest":
            my = H // 2
            arms = [g[1:my, mid], g[my + 1:H - 1, mid], g[my, 1:mid], g[my, mid + 1:W - 1]]
            if all((a != WALL).all() for a in arms):
                n_fallback += 1
                continue
            assert g[my, mid] == WALL and all(int((a != WALL).sum()) == 1 for a in arms)
            off_cross = inner_walls - sum(int((a == WALL).sum()) for a in arms) - 1
            assert 0 <= off_cross <= 12
    assert n_fallback <= L // 50


def test_deterministic_streams_and_addressing():
    a = _env(4, 16, "mediumhard", 7, 300).layouts_numpy()
    b = _env(4, 16, "mediumhard", 7, 300).layouts_numpy()
    c = _env(4, 16, "mediumhard", 8, 300).layouts_numpy()
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert not np.array_equal(a[0], c[0])
    env = _env(4, 16, "mediumhard", 7, 100)
    env.generate_layouts("mediumhard", 7, 100, first_number=200)  # in place: same pool size
    d = env.layouts_numpy()
    assert np.array_equal(d[0], a[0][200:300]) and np.array_equal(d[1], a[1][200:300])
    assert len({cells.tobytes() for cells in a[0]}) >= 299  # layouts of one stream are distinct
    with pytest.raises(ValueError):
        env.generate_layouts("nightmare", 0, 10)
    with pytest.raises(ValueError):
        _env(4, 5, "hardest", 0, 10)  # grid too small for the four-rooms routine


def test_statistics_match_the_host_generator():
    """Same distributions as merlin_b200.layouts (pinned to the reference by fixtures): wall counts, poses, goal
    distances of mediumhard 16x16 agree within sampling error."""
    from merlin_b200 import layouts
    Ld, Lh = 16384, 3000
    dc, da = _env(4, 16, "mediumhard", 99, Ld).layouts_numpy()
    hc, ha = layouts.generate("mediumhard", 16, range(500000, 500000 + Lh))

    def stats(cells, agent):
        g = cells.reshape(-1, 16, 16)
        walls = (g[:, 1:-1, 1:-1] == WALL).sum((1, 2)).astype(np.float64)
        goal = np.array([np.argwhere(x == GOAL)[0] for x in g])  # (y, x)
        dist = np.abs(goal[:, 1] - agent[:, 0]) + np.abs(goal[:, 0] - agent[:, 1])
        return walls, agent[:, 0].astype(np.float64), agent[:, 1].astype(np.float64), dist.astype(np.float64), agent[:, 2]

    d, h = stats(dc, da), stats(hc, ha)
    for k, name in enumerate(["interior walls", "agent x", "agent y", "agent-goal distance"]):
        se = np.sqrt(d[k].var() / len(d[k]) + h[k].var() / len(h[k]))
        assert abs(d[k].mean() - h[k].mean()) < 5 * se, (name, d[k].mean(), h[k].mean(), se)
        assert abs(d[k].std() - h[k].std()) < 0.1 * h[k].std() + 0.05, (name, d[k].std(), h[k].std())
    assert set(np.unique(d[0])) == set(range(19, 40))  # n = integers(19, 40) walls, every value occurs
    counts = np.bincount(d[4], minlength=4) / Ld
    assert np.all(np.abs(counts - 0.25) < 0.02)
    xs = np.bincount(da[:, 0], minlength=16)[1:15] / Ld  # agent column ~ uniform over the 14 interior columns
    assert np.all(np.abs(xs - 1 / 14) < 0.012)


@pytest.mark.parametrize("difficulty,size,N", [("mediumhard", 16, 5000), ("hard", 32, 300), ("hardest", 16, 40)])
def test_rollouts_over_device_generated_layouts_match_the_oracle(difficulty, size, N):
    from merlin_b200 import codes
    L = 777
    env = _env(N, size, difficulty, 2024, L, max_steps=19)
    cells, agent = env.layouts_numpy()
    ref = fast.OracleVecEnv(N, codes.unpack_to_encoding(cells, size, size), agent, max_steps=19)
    obs, sym = env.reset()
    robs, rsym = ref.reset()
    assert np.array_equal(obs.cpu().numpy(), robs) and np.array_equal(sym.cpu().numpy(), rsym)
    rng = np.random.default_rng(0)
    for t in range(45):
        a = rng.integers(0, 3, N)
        obs, r, te, tr, info = env.step(torch.as_tensor(a, device="cuda:0"))
        robs, rr, rte, rtr, _ = ref.step(a)
        assert np.array_equal(r.cpu().numpy(), rr) and np.array_equal(te.cpu().numpy(), rte)
        assert np.array_equal(tr.cpu().numpy(), rtr) and np.array_equal(obs.cpu().numpy(), robs), t
    assert np.array_equal(env.cells_numpy(), cells[env.state_numpy()["layout"]])


def test_ppo_on_device_generated_layouts():
    from src.ppo import PPO
    from src.scenario_creator.scenario_creator import ScenarioCreator
    torch.manual_seed(0)
    env = ScenarioCreator().create_batched_env("mediumhard", 64, device="cuda:0", layouts="device", seeds=5,
                                               n_layouts=512, want_symbolic=True)
    assert env.n_layouts == 512
    agent = PPO(env, batch_size=64 * 8, minibatch_size=128, update_epochs=1, obs_storage="symbolic")
    m = agent.update(agent.collect_rollouts())
    assert np.isfinite(m["pi_loss"]) and np.isfinite(m["v_loss"])
