"""CPU tier, only where the reference checkout is mounted (/root/reference; absent on the GPU box): the REAL reference
modules, imported over oracle/shim.py, against the product's host logic and the kernel arithmetic compiled for the host
-- live, not through fixtures.  (minigrid/gymnasium underneath the reference are the restatement; see oracle/shim.py.)"""
from __future__ import annotations

import os
import subprocess
import sys

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
REF = os.environ.get("MERLIN_REFERENCE_ROOT", "/root/reference")

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "src")), reason="reference checkout not mounted")

# The reference's package is called `src`, like the product's drop-in mirror: run it in a child interpreter so the two
# never share sys.modules.
CHILD = r'''
import sys, json
sys.path[:0] = [REPO_ROOT_DIR, REPO_ROOT_DIR + "/ppo-2dgrid_b200", REPO_ROOT_DIR + "/tests"]
import numpy as np
from oracle import shim
shim.install()
import src.custom_envs.register                                     # the reference's own registration
from src.scenario_creator.scenario_creator import ScenarioCreator   # reference
from src.wrappers.stuck_penalty_wrapper import StuckPenaltyWrapper  # reference
import src.ppo as ref_ppo
assert src.ppo.__file__.startswith(shim.REFERENCE_ROOT), src.ppo.__file__
from merlin_b200 import codes, layouts
import test_host_logic as thl

sc = ScenarioCreator(shim.REFERENCE_ROOT + "/src/config/scenario.yaml")
out = {"layouts": 0, "steps": 0}
rng = np.random.default_rng(0)
for diff in ("easy", "medium", "mediumhard", "hard", "hardest"):
    env = StuckPenaltyWrapper(sc.create_env(diff))
    for seed in (3, 777, 200001):
        obs, _ = env.reset(seed=seed)
        u = env.unwrapped
        # 1. layouts: product generator == the reference's _gen_grid for the same seed
        cells, agent = layouts.generate(diff, 16, [seed])
        assert np.array_equal(codes.unpack_to_encoding(cells, 16, 16)[0], u.grid.encode()), (diff, seed)
        assert tuple(agent[0]) == (u.agent_pos[0], u.agent_pos[1], u.agent_dir), (diff, seed)
        out["layouts"] += 1
        # 2. step + observation + StuckPenalty: kernel arithmetic (host build) == the reference wrapper stack
        hm = thl.HostModelEnv(cells, agent, 16, 16, u.max_steps, 3, True, 0.0)
        rgb, sym, *_ = hm.call(np.zeros(1, np.int64), do_step=False)
        assert np.array_equal(rgb[0], obs)
        for t in range(60):
            a = int(rng.integers(0, 3)) if t % 7 else 2
            obs, r, te, tr, info = env.step(a)
            rgb, sym, rew, hte, htr, hsk = hm.call(np.array([a]), do_step=True)
            assert np.array_equal(rgb[0], obs), (diff, seed, t)
            assert np.array_equal(sym[0], u.gen_obs()["image"]), (diff, seed, t)
            assert rew[0] == np.float32(r) and bool(hte[0]) == te and bool(htr[0]) == tr and bool(hsk[0]) == info["stuck"]
            out["steps"] += 1
            if te or tr:
                break
# 3. random (difficulty, size, seed) triples, including continued streams (reset() without a seed, as PPO training does)
cases = 0
for k in range(60):
    diff = ("easy", "medium", "mediumhard", "hard", "hardest")[int(rng.integers(0, 5))]
    size = int(rng.integers(8, 27))
    seed = int(rng.integers(0, 2**31 - 1))
    sc2 = ScenarioCreator(shim.REFERENCE_ROOT + "/src/config/scenario.yaml")
    sc2.config["difficulties"][diff]["params"]["size"] = size
    env = sc2.create_env(diff)
    env.reset(seed=seed)
    stream = [env.unwrapped]
    cells, agent = layouts.generate_stream(diff, size, seed, 3)
    for j in range(3):
        u = env.unwrapped
        assert np.array_equal(codes.unpack_to_encoding(cells[j:j + 1], size, size)[0], u.grid.encode()), (diff, size, seed, j)
        assert tuple(agent[j]) == (u.agent_pos[0], u.agent_pos[1], u.agent_dir), (diff, size, seed, j)
        env.reset()
    cases += 1
out["random_layout_cases"] = cases
print(json.dumps(out))
'''.replace("REPO_ROOT_DIR", repr(ROOT))


def test_real_reference_modules_agree_with_product_logic():
    res = subprocess.run([sys.executable, "-c", CHILD], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stderr[-3000:]
    import json
    out = json.loads(res.stdout.strip().splitlines()[-1])
    assert out["layouts"] == 15 and out["steps"] > 300 and out["random_layout_cases"] == 60
