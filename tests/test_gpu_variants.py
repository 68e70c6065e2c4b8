"""Parity of the selectable step-kernel mappings that were measured and not adopted as defaults: the CTA-tile kernel
with the frame phase routed through the TMA unit (choice 4: frames assembled in shared memory, one 9408-byte
cp.async.bulk per frame), the group kernel with groups handed out in order (choice 6) and the kernel that serves four
envs per warp, eight lanes each (choice 7: steps of the three-action configuration; everything else asked of it runs
the warp-per-env kernel).  Same bar as every other
mapping: bit-exact observations, rewards, flags, poses and grids against the oracle and the reference fixtures."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import helpers  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=[(4, 0), (6, 0), (7, 0), (3, 2), (6, 2)],
                ids=["tile_tma_kernel", "ordered_kernel", "quad_kernel", "tile_kernel_row_obs", "ordered_kernel_row_obs"])
def kernel_choice(request):
    """(kernel choice, observation path): path 2 = the row-parallel gen_obs (obs_swar.cuh) in kernels that default to the
    per-cell form (the symbolic-only kernel uses it by default and is covered by the main suites)."""
    from merlin_b200 import set_kernel_choice, set_observation_path
    set_kernel_choice(request.param[0])
    set_observation_path(request.param[1])
    yield request.param
    set_kernel_choice(0)
    set_observation_path(0)


def _parity():
    import test_gpu_parity as tp
    return tp


@pytest.mark.parametrize("name", helpers.trace_names())
def test_reference_traces(name):
    tp = _parity()
    helpers.replay_trace_autoreset(tp._make_gpu_env, helpers.load(f"trace_{name}.npz"))
    helpers.replay_trace_manual_reset(tp._make_gpu_env, helpers.load(f"trace_{name}.npz"))


@pytest.mark.parametrize("N", [1, 31, 1000, 40000])
def test_random_rollouts_vs_oracle(N):
    tp = _parity()
    _, codes, _, layouts, _ = tp._mods()
    cells, agent = layouts.generate("mediumhard", 16, range(500, 564))
    enc = codes.unpack_to_encoding(cells, 16, 16)
    tp._compare_batched(N, enc, agent, 60 if N < 2000 else 12, max_steps=25, seed=N)
    tp._compare_batched(N, enc, agent, 30 if N < 2000 else 8, max_steps=25, seed=N + 1, stuck_penalty=True,
                        exploration_bonus=0.01)


@pytest.mark.parametrize("N", [2, 5, 6, 7, 4097, 6000, 20000])
def test_ragged_quads_and_many_restarts(N):
    """Batch sizes that leave the last group of four envs ragged, one-round and multi-round launch shapes, short episodes
    (most steps restart some env onto another layout)."""
    tp = _parity()
    _, codes, _, layouts, _ = tp._mods()
    cells, agent = layouts.generate("mediumhard", 16, range(5000, 5257))
    enc = codes.unpack_to_encoding(cells, 16, 16)
    _, _, n_done = tp._compare_batched(N, enc, agent, 40 if N < 5000 else 16, max_steps=7, seed=N,
                                       check_every=1 if N < 5000 else 5)
    assert n_done > 0


@pytest.mark.parametrize("W,H,N", [(3, 3, 41), (5, 12, 97), (19, 7, 97), (255, 255, 66), (255, 4, 65)])
def test_grid_extents(W, H, N):
    """3x3 (every window mostly outside the grid) to 255x255, W != H; doors and keys in a three-action pool (opaque
    closed / locked doors, Grid.encode's state byte)."""
    tp = _parity()
    rng = np.random.default_rng(W * 1000 + H)
    enc, agent = tp._rect_layouts(rng, 24 if W * H < 10000 else 6, W, H)
    tp._compare_batched(N, enc, agent, 40, max_steps=13, seed=W + H)
    enc, agent = tp._object_layouts(rng, 48, 11)
    tp._compare_batched(203, enc, agent, 40, max_steps=17, seed=W)


def test_symbolic_only_kernel_per_cell_and_row_forms_agree():
    """The symbolic-only kernel under observation path 1 (per-cell) and 0 / 2 (row-parallel): same images, bit for bit,
    on grids with every object type, ragged batch sizes and a grid narrower than the window (always per-cell)."""
    import torch
    from merlin_b200 import set_kernel_choice, set_observation_path
    tp = _parity()
    rng = np.random.default_rng(17)
    for size, N in ((11, 4099), (16, 33), (6, 500)):
        enc, agent = tp._object_layouts(rng, 64, size)
        outs = []
        for path in (1, 0):
            set_kernel_choice(0)
            set_observation_path(path)
            env = tp._make_gpu_env(N, enc, agent, n_actions=7, max_steps=19, want_rgb=False)
            r2 = np.random.default_rng(5)
            _, sym = env.reset()
            frames = [sym.clone()]
            for _ in range(25):
                _, _, _, _, info = env.step(torch.as_tensor(r2.integers(0, 7, N), device="cuda:0"))
                frames.append(info["obs_symbolic"].clone())
            outs.append(torch.stack(frames))
        assert torch.equal(outs[0], outs[1]), size


def test_seven_actions_objects_and_guard_bands():
    tp = _parity()
    rng = np.random.default_rng(3)
    enc, agent = tp._object_layouts(rng, 96, 11)
    _, _, n_done = tp._compare_batched(3000, enc, agent, 40, n_actions=7, max_steps=23, seed=9)
    assert n_done > 0
    tp.test_outputs_stay_inside_their_buffers(257)
    tp.test_outputs_stay_inside_their_buffers(9473)
    tp.test_masked_reset_only_touches_selected_envs()
