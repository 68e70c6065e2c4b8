"""CPU tests of the C ABI boundary: the shared library loads, exports every symbol include/merlin_b200.h
declares, and -- with no CUDA device -- fails loudly instead of falling back to a CPU path."""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge._load_build_module().build_library()
    from merlin_b200 import _lib
    return _lib.load()


def test_header_symbols_are_exported(lib):
    header = open(os.path.join(ROOT, "include", "merlin_b200.h")).read()
    declared = set(re.findall(r"\b(merlin_[a-z_0-9]+)\s*\(", header))
    assert len(declared) >= 15
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in merlin_b200.h but not exported"
    from merlin_b200 import _lib
    assert declared == set(_lib.EXPORTS)


def test_pack_cell_matches_python_codes(lib):
    from merlin_b200 import codes
    for t in range(1, 10):
        for c in range(6):
            for s in range(3):
                assert lib.merlin_pack_cell(t, c, s) == int(codes.pack(t, c, s)), (t, c, s)


def test_config_struct_layout(lib):
    from merlin_b200 import _lib
    cfg = _lib.EnvConfig()
    lib.merlin_env_default_config(C.byref(cfg))
    assert (cfg.n_envs, cfg.width, cfg.height, cfg.view, cfg.tile) == (1, 16, 16, 7, 8)
    assert cfg.flags == _lib.F_AUTO_RESET and cfg.stuck_max_stay == 3
    assert cfg.stuck_penalty == -0.1 and cfg.explore_bonus == 0.0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-device behaviour")
def test_no_cpu_fallback(lib):
    from merlin_b200 import BatchedMerlinEnv, _lib, gae, layouts
    cfg = _lib.EnvConfig()
    lib.merlin_env_default_config(C.byref(cfg))
    h = C.c_void_p()
    assert lib.merlin_env_create(C.byref(cfg), C.byref(h)) == _lib.ECUDA
    assert b"no CPU fallback" in lib.merlin_last_error()
    cells, agent = layouts.generate("medium", 16, [0])
    with pytest.raises(RuntimeError):
        BatchedMerlinEnv(1, cells, agent, width=16, height=16, device="cpu")
    with pytest.raises(RuntimeError):
        gae(torch.zeros(4), torch.zeros(4), torch.zeros(4), 0.0)
    x = np.zeros(4, np.float32)
    rc = lib.merlin_gae(x.ctypes.data, x.ctypes.data, x.ctypes.data, x.ctypes.data, x.ctypes.data, x.ctypes.data,
                        4, 1, 0.99, 0.95, None)
    assert rc == _lib.ECUDA


def test_tuning_knobs_validate_their_range(lib):
    """The process-wide defaults of the two knobs need no device: every documented value is accepted, anything else is
    MERLIN_EINVAL with a message, and the defaults end where they started (automatic)."""
    from merlin_b200 import _lib
    try:
        for choice in range(0, 8):   # 0 automatic .. 7 four envs per warp (include/merlin_b200.h)
            assert lib.merlin_set_kernel_choice(choice) == 0, choice
        for bad in (-1, 8, 99):
            assert lib.merlin_set_kernel_choice(bad) == _lib.EINVAL, bad
            assert b"kernel choice" in lib.merlin_last_error()
        for path in (0, 1, 2):
            assert lib.merlin_set_observation_path(path) == 0, path
        assert lib.merlin_set_observation_path(3) == _lib.EINVAL
    finally:
        lib.merlin_set_kernel_choice(0)
        lib.merlin_set_observation_path(0)
