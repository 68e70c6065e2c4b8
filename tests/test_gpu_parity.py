"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI
(libmerlin_b200.so via merlin_b200.BatchedMerlinEnv / merlin_b200.gae), against the oracle.

Bars: observations (RGB + symbolic), rewards, terminated/truncated flags, poses and grids are BIT-EXACT;
GAE advantages/returns are bit-exact on the reference fixtures and within 1e-6 relative elsewhere
(north_star tolerance)."""
from __future__ import annotations

import numpy as np
import pytest
import torch

import helpers
from oracle import fast

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=[1, 2, 3], ids=["group_kernel", "warp_kernel", "tile_kernel"])
def kernel_choice(request):
    """Every parity test runs against all three step kernels (warp owns a group / warp per env / CTA tile); the three
    measured-and-not-adopted mappings (TMA frame stores, ordered groups, four envs per warp) have tests/test_gpu_variants.py."""
    from merlin_b200 import set_kernel_choice
    set_kernel_choice(request.param)
    yield request.param
    set_kernel_choice(0)


def _mods():
    from merlin_b200 import BatchedMerlinEnv, codes, gae, layouts, tiles
    return BatchedMerlinEnv, codes, gae, layouts, tiles


def _make_gpu_env(num_envs, enc, agent, **kw):
    BatchedMerlinEnv = _mods()[0]
    return BatchedMerlinEnv(num_envs, enc=enc, agent=agent, device="cuda:0", **kw)


def _np(x):
    return x.detach().cpu().numpy()


# ---- golden traces (produced by the reference's own code) ---------------------------------------------
@pytest.mark.parametrize("name", helpers.trace_names())
def test_trace_autoreset(name):
    helpers.replay_trace_autoreset(_make_gpu_env, helpers.load(f"trace_{name}.npz"))


@pytest.mark.parametrize("name", helpers.trace_names())
def test_trace_manual_reset(name):
    helpers.replay_trace_manual_reset(_make_gpu_env, helpers.load(f"trace_{name}.npz"))


# ---- random batches vs the C oracle ---------------------------------------------------------------------
def _compare_batched(N, enc, agent, steps, n_actions=3, seed=0, check_every=1, **kw):
    env = _make_gpu_env(N, enc, agent, n_actions=n_actions, **kw)
    okw = dict(kw)
    ref = fast.OracleVecEnv(N, enc, agent, n_actions=n_actions, **okw)
    obs, sym = env.reset()
    robs, rsym = ref.reset()
    assert np.array_equal(_np(obs), robs) and np.array_equal(_np(sym), rsym)
    rng = np.random.default_rng(seed)
    n_done = 0
    for t in range(steps):
        a = rng.integers(0, n_actions, N)
        if n_actions == 7:
            a = np.where(rng.random(N) < 0.4, 2, a)
        obs, r, te, tr, info = env.step(torch.as_tensor(a, device="cuda:0"))
        robs, rr, rte, rtr, rinfo = ref.step(a)
        assert np.array_equal(_np(r), rr), t
        assert np.array_equal(_np(te), rte) and np.array_equal(_np(tr), rtr), t
        assert np.array_equal(_np(info["episode_length"]), rinfo["episode_length"]), t
        assert np.array_equal(_np(info["episode_return"]), rinfo["episode_return"]), t
        assert np.array_equal(_np(info["stuck"]), rinfo["stuck"]), t
        n_done += int((rte | rtr).sum())
        if t % check_every == 0 or t == steps - 1:
            assert np.array_equal(_np(info["obs_symbolic"]), rinfo["obs_symbolic"]), t
            assert np.array_equal(_np(obs), robs), t
    pose = env.pose_numpy()
    assert np.array_equal(pose, helpers.get_pose(ref))
    return env, ref, n_done


@pytest.mark.parametrize("N", [1, 31, 33, 37, 4096, 6000, 10000, 20000, 40000])
def test_random_rollout_all_group_sizes(N):
    """Covers every kernel instantiation (groups of 4/8/16/32, tiles of 8/16/32), ragged tails and many auto-resets."""
    _, codes, _, layouts, _ = _mods()
    L = 257
    cells, agent = layouts.generate("mediumhard", 16, range(5000, 5000 + L))
    enc = codes.unpack_to_encoding(cells, 16, 16)
    steps = 60 if N <= 4096 else 24
    _, _, n_done = _compare_batched(N, enc, agent, steps, max_steps=11, seed=N, check_every=1 if N <= 4096 else 6)
    assert n_done > 0


@pytest.mark.parametrize("diff,size", [("easy", 16), ("medium", 16), ("hard", 16), ("hardest", 16), ("hard", 32),
                                       ("mediumhard", 8), ("hardest", 24)])
def test_random_rollout_other_difficulties_and_sizes(diff, size):
    _, codes, _, layouts, _ = _mods()
    cells, agent = layouts.generate(diff, size, range(100))
    enc = codes.unpack_to_encoding(cells, size, size)
    _compare_batched(512, enc, agent, 50, max_steps=17, seed=size)


def _rect_layouts(rng, L, W, H, wall_p=0.15):
    """Random W x H rooms (W != H allowed): border walls, random interior walls, one goal when there is room."""
    enc = np.zeros((L, W, H, 3), np.uint8)
    enc[..., 0] = 1
    enc[:, 0, :, :] = enc[:, -1, :, :] = enc[:, :, 0, :] = enc[:, :, -1, :] = (2, 5, 0)
    agent = np.zeros((L, 3), np.int32)
    for l in range(L):
        inner = rng.random((W - 2, H - 2)) < wall_p
        enc[l, 1:-1, 1:-1][inner] = (2, 5, 0)
        free = np.argwhere(enc[l, :, :, 0] == 1)
        if len(free) == 0:
            enc[l, 1, 1] = (1, 0, 0)
            free = np.array([[1, 1]])
        k = rng.permutation(len(free))
        agent[l] = (free[k[0]][0], free[k[0]][1], rng.integers(0, 4))
        if len(free) > 1:
            enc[l, free[k[1]][0], free[k[1]][1]] = (8, 1, 0)
    return enc, agent


@pytest.mark.parametrize("W,H,N", [(3, 3, 40), (5, 12, 97), (19, 7, 97), (64, 64, 333), (255, 255, 65), (255, 4, 65)])
def test_non_square_minimum_and_maximum_grid_sizes(W, H, N):
    """Grid extents from the 3x3 minimum to the 255x255 maximum of the packed pose, W != H included."""
    rng = np.random.default_rng(W * 1000 + H)
    enc, agent = _rect_layouts(rng, 24 if W * H < 10000 else 6, W, H)
    _compare_batched(N, enc, agent, 40, max_steps=13, seed=W + H)
    _compare_batched(N, enc, agent, 12, n_actions=7, max_steps=9, seed=W)


def _object_layouts(rng, L, size):
    enc = np.zeros((L, size, size, 3), np.uint8)
    enc[..., 0] = 1
    enc[:, 0, :, :] = enc[:, -1, :, :] = enc[:, :, 0, :] = enc[:, :, -1, :] = (2, 5, 0)
    agent = np.zeros((L, 3), np.int32)
    objs = [(2, 5, 0), (2, 5, 0), (8, 1, 0), (9, 0, 0), (3, 2, 0), (4, 4, 0), (4, 4, 1), (4, 4, 2), (5, 4, 0),
            (6, 0, 0), (7, 3, 0), (4, 2, 1), (5, 2, 0)]
    for l in range(L):
        for _ in range(int(rng.integers(5, 30))):
            x, y = rng.integers(1, size - 1, 2)
            enc[l, x, y] = objs[int(rng.integers(0, len(objs)))]
        while True:
            x, y = rng.integers(1, size - 1, 2)
            if enc[l, x, y, 0] == 1:
                break
        agent[l] = (x, y, rng.integers(0, 4))
    return enc, agent


@pytest.mark.parametrize("N", [300, 40000])
def test_seven_actions_full_object_set_mutable_grids(N):
    """pickup / drop / toggle with doors, keys, balls, boxes, lava, floor; grids mutate and are restored on restart."""
    _, codes, _, _, _ = _mods()
    rng = np.random.default_rng(3)
    enc, agent = _object_layouts(rng, 96, 11)
    env, ref, n_done = _compare_batched(N, enc, agent, 70 if N < 1000 else 20, n_actions=7, max_steps=23, seed=N,
                                        check_every=1 if N < 1000 else 5)
    assert n_done > 0
    got = codes.unpack_to_encoding(env.cells_numpy(), 11, 11)
    want = np.stack([ref.gt, ref.gc, ref.gs], axis=-1).reshape(N, 11, 11, 3).transpose(0, 2, 1, 3)
    assert np.array_equal(got, want)
    assert np.array_equal(env.state_numpy()["carry"] & 0xF, ref.carry_t)


def test_wrappers_stuck_and_exploration_bonus():
    _, codes, _, layouts, _ = _mods()
    cells, agent = layouts.generate("hardest", 16, range(64))
    enc = codes.unpack_to_encoding(cells, 16, 16)
    _compare_batched(2048, enc, agent, 80, max_steps=40, stuck_penalty=True, seed=1)
    _compare_batched(2048, enc, agent, 80, max_steps=40, exploration_bonus=0.01, seed=2)
    _compare_batched(2048, enc, agent, 80, max_steps=40, stuck_penalty=True, exploration_bonus=0.02,
                     stuck_max_stay=2, stuck_penalty_value=-0.25, seed=3)


def test_reset_mode_same_and_cursors_fomaml():
    """FOMAML: every episode of a task replays the same layout (reference src/fomaml.py:63,92)."""
    _, codes, _, layouts, _ = _mods()
    cells, agent = layouts.generate("mediumhard", 16, range(40))
    enc = codes.unpack_to_encoding(cells, 16, 16)
    env, ref, _ = _compare_batched(64, enc, agent, 50, max_steps=9, reset_mode="same", seed=5)
    assert np.array_equal(env.state_numpy()["layout"], np.arange(64) % 40)
    env.set_cursors(np.full(64, 7, np.int32))
    obs, sym = env.reset()
    assert np.all(env.state_numpy()["layout"] == 7)
    assert all(np.array_equal(_np(obs)[0], _np(obs)[k]) for k in range(64))


def test_masked_reset_only_touches_selected_envs():
    _, codes, _, layouts, _ = _mods()
    cells, agent = layouts.generate("medium", 16, range(50))
    enc = codes.unpack_to_encoding(cells, 16, 16)
    N = 100
    env = _make_gpu_env(N, enc, agent, auto_reset=False)
    ref = fast.OracleVecEnv(N, enc, agent, auto_reset=False)
    env.reset(); ref.reset()
    rng = np.random.default_rng(0)
    for _ in range(5):
        a = rng.integers(0, 3, N)
        env.step(torch.as_tensor(a, device="cuda:0")); ref.step(a)
    mask = rng.random(N) < 0.3
    before = _np(env.obs).copy()
    obs, sym = env.reset(torch.as_tensor(mask, device="cuda:0"))
    robs, rsym = ref.reset(mask)
    assert np.array_equal(_np(obs)[mask], robs[mask]) and np.array_equal(_np(sym)[mask], rsym[mask])
    assert np.array_equal(_np(obs)[~mask], before[~mask])  # untouched rows
    assert np.array_equal(env.pose_numpy(), helpers.get_pose(ref))


# ---- GAE --------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tag", ["ppo_2048", "ppo_256", "fomaml_256", "alldone_64", "t1"])
def test_gae_bit_exact_on_reference_fixtures(tag):
    gae = _mods()[2]
    fx = helpers.load("gae.npz")
    dev = "cuda:0"
    adv, ret = gae(torch.tensor(fx[f"{tag}_rew"], device=dev), torch.tensor(fx[f"{tag}_val"], device=dev),
                   torch.tensor(fx[f"{tag}_done"], device=dev), float(fx[f"{tag}_last"]),
                   float(fx[f"{tag}_gamma"]), float(fx[f"{tag}_lam"]))
    assert np.array_equal(_np(adv), fx[f"{tag}_adv_torch"]) and np.array_equal(_np(adv), fx[f"{tag}_adv_numpy"])
    assert np.array_equal(_np(ret), fx[f"{tag}_ret_torch"])


@pytest.mark.parametrize("T,N", [(128, 4096), (1, 5), (7, 33), (2048, 1), (9, 130), (15, 4097), (16, 5), (23, 75777),
                                 (300, 200), (129, 128), (256, 8192), (17, 8193)])   # tile kernel: several windows, range ends
def test_gae_batched_vs_oracle(T, N):
    gae = _mods()[2]
    g = torch.Generator(device="cpu").manual_seed(T * 7 + N)
    rew = torch.rand(T, N, generator=g) * (torch.rand(T, N, generator=g) < 0.05)
    val = torch.randn(T, N, generator=g)
    done = (torch.rand(T, N, generator=g) < 0.01).float()
    last = torch.randn(N, generator=g)
    adv, ret = gae(rew.cuda(), val.cuda(), done.cuda(), last.cuda(), 0.995, 0.95)
    radv, rret = fast.gae(rew.numpy(), val.numpy(), done.numpy(), last.numpy(), 0.995, 0.95)
    # north_star tolerance: 1e-6 relative in fp32 (the kernel is in fact bit-identical to the unfused loop)
    np.testing.assert_allclose(_np(adv), radv, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(_np(ret), rret, rtol=1e-6, atol=1e-7)
    assert np.array_equal(_np(adv), radv) and np.array_equal(_np(ret), rret)


# ---- size-independent properties at BASELINE scale --------------------------------------------------------
def test_full_scale_properties_262144_envs():
    """262 144 envs x 16x16 (config 2 scale): determinism, RGB == atlas expansion of the symbolic view,
    episode lengths bounded, flags consistent, no bad actions."""
    BatchedMerlinEnv, codes, _, layouts, tiles = _mods()
    N, L = 262144, 1024
    cells, agent = layouts.generate("mediumhard", 16, range(9000, 9000 + L))
    dev = "cuda:0"
    atlas = torch.as_tensor(tiles.build_atlas(), device=dev)  # [128, 8, 8, 3]

    def expand(sym):  # RGB frame implied by the symbolic view: tile = atlas[packed code] (0 when unseen)
        t, c = sym[..., 0].long(), sym[..., 1].long()
        kind = torch.where(t == 0, torch.zeros_like(t), t | (c << 4))
        kind[:, 3, 6] = 10  # agent cell
        tiles_ = atlas[kind]  # [N, vi, vj, py, px, 3]
        return tiles_.permute(0, 2, 3, 1, 4, 5).reshape(sym.shape[0], 56, 56, 3)

    outs = []
    for rep in range(2):
        env = BatchedMerlinEnv(N, cells, agent, width=16, height=16, max_steps=64, device=dev)
        obs, sym = env.reset()
        g = torch.Generator(device=dev).manual_seed(123)
        lens = torch.zeros(N, dtype=torch.int64, device=dev)
        acc = torch.zeros((), dtype=torch.int64, device=dev)
        for t in range(130):
            a = torch.randint(0, 3, (N,), generator=g, device=dev)
            obs, r, te, tr, info = env.step(a)
            done = te | tr
            assert bool(((info["episode_length"] > 0) == done).all())
            assert int(info["episode_length"].max()) <= 64
            assert bool((r[~te] == 0).all()) and bool((r[te] > 0).all())
            acc += obs[::97].long().sum() + info["obs_symbolic"].long().sum() * 31 + (r * 1e4).long().sum()
            if t % 43 == 0:
                for lo in range(0, N, 65536):
                    assert torch.equal(obs[lo:lo + 65536], expand(info["obs_symbolic"][lo:lo + 65536]))
        assert env.bad_actions() == 0
        st = env.state_numpy()
        assert st["step_count"].max() < 64 and st["step_count"].min() >= 0
        outs.append((int(acc), obs.clone()))
        env.close()
    assert outs[0][0] == outs[1][0] and torch.equal(outs[0][1], outs[1][1])


# ---- API behaviour ----------------------------------------------------------------------------------------
def test_errors_and_bad_actions():
    BatchedMerlinEnv, codes, _, layouts, _ = _mods()
    cells, agent = layouts.generate("medium", 16, range(4))
    env = BatchedMerlinEnv(8, cells, agent, width=16, height=16, device="cuda:0")
    with pytest.raises(RuntimeError):
        env.step(torch.zeros(8, dtype=torch.int64, device="cuda:0"))  # step before reset
    env.reset()
    before = env.pose_numpy()
    env.step(torch.tensor([0, 1, 2, 3, 7, -1, 2, 99], device="cuda:0"))  # 3,7,-1,99 are out of the 3-action range
    assert env.bad_actions() == 4
    after = env.pose_numpy()
    assert np.array_equal(after[[3, 4, 5, 7], :3], before[[3, 4, 5, 7], :3])  # executed as no-ops
    assert np.all(after[:, 3] == 1)
    with pytest.raises(ValueError):
        env.step(torch.zeros(5, dtype=torch.int64, device="cuda:0"))
    bad = cells.copy(); bad[0, 5] = 0x0A  # 'agent' is not a grid object
    with pytest.raises(ValueError):
        BatchedMerlinEnv(8, bad, agent, width=16, height=16, device="cuda:0")
    with pytest.raises(RuntimeError):
        BatchedMerlinEnv(8, cells, agent, width=16, height=16, device="cpu")


def test_step_writes_into_rollout_slot_and_cuda_graph():
    """out_obs lets step() write straight into a [T, N, 56, 56, 3] rollout buffer; step is graph-capturable."""
    BatchedMerlinEnv, codes, _, layouts, _ = _mods()
    cells, agent = layouts.generate("mediumhard", 16, range(32))
    enc = codes.unpack_to_encoding(cells, 16, 16)
    N, T = 512, 6
    dev = "cuda:0"
    env = BatchedMerlinEnv(N, cells, agent, width=16, height=16, max_steps=5, device=dev)
    ref = fast.OracleVecEnv(N, enc, agent, max_steps=5)
    buf = torch.zeros((T + 1, N, 56, 56, 3), dtype=torch.uint8, device=dev)
    env.reset(out_obs=buf[0]); ref.reset()
    acts = torch.randint(0, 3, (T, N), device=dev)
    for t in range(T):
        env.step(acts[t], out_obs=buf[t + 1])
        robs, *_ = ref.step(_np(acts[t]))
        assert np.array_equal(_np(buf[t + 1]), robs)
    # CUDA graph: capture one step reading a static action tensor, replay it T times
    env2 = BatchedMerlinEnv(N, cells, agent, width=16, height=16, max_steps=5, device=dev)
    ref2 = fast.OracleVecEnv(N, enc, agent, max_steps=5)
    env2.reset(); ref2.reset()
    static_a = torch.zeros(N, dtype=torch.int64, device=dev)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        static_a.copy_(acts[0]); env2.step(static_a); ref2.step(_np(acts[0]))  # warm-up outside capture
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        env2.step(static_a)
    for t in range(1, T):
        static_a.copy_(acts[t])
        graph.replay()
        robs, rr, *_ = ref2.step(_np(acts[t]))
        assert np.array_equal(_np(env2.obs), robs) and np.array_equal(_np(env2.reward), rr)


# ---- frames from stored symbolic observations (merlin_env_render) -----------------------------------------
def _blocked_of(rgb):
    """Reference layout transform on the host: u8[M,56,56,3] -> u8[M,14,14,48], channel index c*16 + dy*4 + dx."""
    m = rgb.shape[0]
    return rgb.reshape(m, 14, 4, 14, 4, 3).transpose(0, 1, 3, 5, 2, 4).reshape(m, 14, 14, 48)


@pytest.mark.parametrize("n_actions", [3, 7])
def test_render_from_symbolic_is_bit_identical_to_step_frames(n_actions):
    """RGBImgPartialObsWrapper.observation on STORED symbolic images == the frames step() wrote, with and without a
    row gather, in the reference layout and in the blocked (space-to-depth) layout; objects / carrying included."""
    _, codes, _, layouts, _ = _mods()
    rng = np.random.default_rng(11)
    N, T = 300, 24
    if n_actions == 3:
        cells, agent = layouts.generate("hardest", 16, range(64))
        enc = codes.unpack_to_encoding(cells, 16, 16)
    else:
        enc, agent = _object_layouts(rng, 64, 11)
    env = _make_gpu_env(N, enc, agent, n_actions=n_actions, max_steps=15)
    sym_store = torch.zeros((T + 1, N, 7, 7, 3), dtype=torch.uint8, device="cuda:0")
    rgb_store = torch.zeros((T + 1, N, 56, 56, 3), dtype=torch.uint8, device="cuda:0")
    env.reset(out_obs=rgb_store[0], out_symbolic=sym_store[0])
    for t in range(T):
        a = rng.integers(0, n_actions, N)
        if n_actions == 7:
            a = np.where(rng.random(N) < 0.3, 3, a)  # plenty of pickups so that carried objects show under the agent
        env.step(torch.as_tensor(a, device="cuda:0"), out_obs=rgb_store[t + 1], out_symbolic=sym_store[t + 1])
    if n_actions == 7:
        assert int((env.state_numpy()["carry"] != 0).sum()) > 0
    flat_rgb = rgb_store.reshape(-1, 56, 56, 3)
    got = env.render(sym_store)  # all rows, reference layout
    assert got.shape == flat_rgb.shape and torch.equal(got, flat_rgb)
    idx = torch.as_tensor(rng.integers(0, flat_rgb.shape[0], 1000), device="cuda:0")  # gather with repeats
    assert torch.equal(env.render(sym_store, idx), flat_rgb[idx])
    blocked = env.render(sym_store, idx, blocked=True)
    assert blocked.shape == (1000, 14, 14, 48)
    assert np.array_equal(_np(blocked), _blocked_of(_np(flat_rgb[idx])))
    out = torch.zeros((7, 56, 56, 3), dtype=torch.uint8, device="cuda:0")
    assert env.render(sym_store, idx[:7], out=out).data_ptr() == out.data_ptr() and torch.equal(out, flat_rgb[idx[:7]])
    with pytest.raises(ValueError):
        env.render(sym_store.float())
    assert env.render(sym_store[:0]).shape == (0, 56, 56, 3)
    # float32 frames (merlin_env_render_f32): the blocked layout as pixel values, or as the `x / 255.0` of the reference's
    # CNNFeatureExtractor.forward (src/actor_critic.py:21) -- bit-identical to torch's CPU kernel (IEEE division) and
    # to torch's CUDA kernel (multiplication by 1/255) respectively
    want_px = torch.as_tensor(_blocked_of(_np(flat_rgb[idx]))).float()
    f_norm = env.render(sym_store, idx, blocked=True, dtype=torch.float32, normalise=True)
    assert f_norm.dtype == torch.float32 and f_norm.shape == (1000, 14, 14, 48)
    assert torch.equal(f_norm.cpu(), want_px / 255.0)
    want_px = want_px.cuda()
    assert torch.equal(env.render(sym_store, idx, blocked=True, dtype=torch.float32, normalise="reciprocal"), want_px / 255.0)
    with pytest.raises(ValueError):
        env.render(sym_store, idx, blocked=True, dtype=torch.float32, normalise="half")
    assert torch.equal(env.render(sym_store, idx, blocked=True, dtype=torch.float32), want_px)
    f_all = env.render(sym_store, blocked=True, dtype=torch.float32, normalise=True)  # all rows, ragged last group
    assert torch.equal(f_all.cpu(), torch.as_tensor(_blocked_of(_np(flat_rgb))).float() / 255.0)
    guard = torch.full((5 * 9408 + 64,), -7.0, device="cuda:0")
    env.render(sym_store, idx[:5], out=guard[32:32 + 5 * 9408].view(5, 14, 14, 48), blocked=True, dtype=torch.float32,
               normalise=True)
    assert bool((guard[:32] == -7.0).all()) and bool((guard[-32:] == -7.0).all())
    assert torch.equal(guard[32:-32].view(5, 14, 14, 48), f_norm[:5])
    with pytest.raises(ValueError):
        env.render(sym_store, dtype=torch.float32)  # float32 exists in the blocked layout only
    assert env.render(sym_store[:0], blocked=True, dtype=torch.float32).shape == (0, 14, 14, 48)


def test_render_f32_replayed_from_a_graph_after_the_pool_gained_tile_kinds():
    """A CUDA graph captured while the layout pool shows 5 tile kinds stays correct after an equal-sized re-upload
    brings kinds the kernel's float atlas was not sized for (they are converted on the fly from the u8 atlas)."""
    _, codes, _, layouts, _ = _mods()
    rng = np.random.default_rng(5)
    cells, agent = layouts.generate("mediumhard", 11, range(32))
    env = _make_gpu_env(64, codes.unpack_to_encoding(cells, 11, 11), agent, max_steps=12)
    sym = torch.zeros((64, 7, 7, 3), dtype=torch.uint8, device="cuda:0")
    rgb = torch.zeros((64, 56, 56, 3), dtype=torch.uint8, device="cuda:0")
    out = torch.zeros((64, 14, 14, 48), dtype=torch.float32, device="cuda:0")
    env.reset(out_obs=rgb, out_symbolic=sym)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        env.render(sym, out=out, blocked=True, dtype=torch.float32, normalise=True)
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        env.render(sym, out=out, blocked=True, dtype=torch.float32, normalise=True)
    g.replay()
    assert torch.equal(out.cpu(), torch.as_tensor(_blocked_of(_np(rgb))).float() / 255.0)
    enc2, agent2 = _object_layouts(rng, 32, 11)  # lava, floor, doors, keys, balls, boxes: 13 more tile kinds
    env.upload_layouts(codes.pack_encoding(enc2), agent2)
    env.reset(out_obs=rgb, out_symbolic=sym)
    for _ in range(6):
        env.step(torch.as_tensor(rng.integers(0, 3, 64), device="cuda:0"), out_obs=rgb, out_symbolic=sym)
    assert len(np.unique(_np(sym)[..., 0])) > 5
    g.replay()
    assert torch.equal(out.cpu(), torch.as_tensor(_blocked_of(_np(rgb))).float() / 255.0)


@pytest.mark.parametrize("N", [1, 33, 257, 5000, 9473])
def test_outputs_stay_inside_their_buffers(N):
    """Guard bands around every caller-owned output (compute-sanitizer is not available on the GPU pool): frames,
    symbolic rows and per-env scalars of a ragged batch are written, the bytes before and after them are not."""
    _, codes, _, layouts, _ = _mods()
    cells, agent = layouts.generate("mediumhard", 16, range(40))
    env = _make_gpu_env(N, codes.unpack_to_encoding(cells, 16, 16), agent, max_steps=7)
    G = 4096  # guard bytes on each side (multiple of 16: the frame pointer must stay 16-byte aligned)
    dev = "cuda:0"

    def guarded(nbytes):
        buf = torch.full((G + nbytes + G,), 0xA5, dtype=torch.uint8, device=dev)
        return buf, buf[G:G + nbytes]

    rgb_all, rgb = guarded(N * 56 * 56 * 3)
    sym_all, sym = guarded(N * 147 + (-N * 147) % 16)
    raw, views = {}, {}
    for name, dtype, width in (("reward", torch.float32, 4), ("terminated", torch.bool, 1), ("truncated", torch.bool, 1),
                               ("episode_return", torch.float32, 4), ("episode_length", torch.int32, 4),
                               ("stuck", torch.bool, 1), ("done", torch.float32, 4)):
        whole, inner = guarded(N * width + (-N * width) % 16)
        raw[name] = whole
        views[name] = inner[: N * width].view(dtype)
    bufs = env.make_step_buffers(**views)  # the step kernel writes straight into these (rollout-row style)
    rgb_v, sym_v = rgb.view(N, 56, 56, 3), sym[: N * 147].view(N, 7, 7, 3)
    env.reset(out_obs=rgb_v, out_symbolic=sym_v)
    rng = np.random.default_rng(N)
    for _ in range(14):  # max_steps 7: the last step ends every env's second episode (unless it reached the goal early)
        env.step(torch.as_tensor(rng.integers(0, 3, N), device=dev), out_obs=rgb_v, out_symbolic=sym_v, out=bufs)
    torch.cuda.synchronize()
    for whole in [rgb_all, sym_all] + list(raw.values()):
        assert bool((whole[:G] == 0xA5).all()) and bool((whole[-G:] == 0xA5).all())
    assert bool((sym_all[G + N * 147: G + sym.numel()] == 0xA5).all())  # padding after the last symbolic row
    assert torch.equal(bufs.done, (bufs.terminated | bufs.truncated).float())
    assert 0 < int(bufs.episode_length.max()) <= 7 and not bool((rgb_v == 0xA5).all(dim=(1, 2, 3)).any())
