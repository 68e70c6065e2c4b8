#!/usr/bin/env python
"""GAE kernel timing: algorithmic bytes (20 B per (t, env) + 4 B per env) / CUDA-event time, against the measured HBM
peak, at the training size (T=128, N=4096: launch-bound) and at streaming sizes (N*T >= 2^26), beside the reference's
Python loop (src/ppo.py:107-120) restated in the oracle, on the CPU.

    python tools/bench_gae.py --out profiles/r01_gae.json
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "ppo-2dgrid_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import torch
    from merlin_b200 import _lib, gae

    dev = torch.device("cuda", 0)
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    rows = []
    for T, N in [(2048, 1), (128, 32), (128, 1024), (128, 4096), (128, 16384), (128, 65536), (128, 524288), (1024, 65536), (64, 1048576), (256, 1048576)]:
        rew = torch.rand(T, N, device=dev)
        val = torch.randn(T, N, device=dev)
        done = (torch.rand(T, N, device=dev) < 0.01).float()
        last = torch.randn(N, device=dev)
        adv, ret = gae(rew, val, done, last)
        lib = _lib.load()
        stream = torch.cuda.current_stream().cuda_stream

        def call():  # the C-ABI entry point on preallocated outputs (the Python wrapper adds ~30 us of host time)
            _lib.check(lib.merlin_gae(rew.data_ptr(), val.data_ptr(), done.data_ptr(), last.data_ptr(), adv.data_ptr(),
                                      ret.data_ptr(), T, N, 0.99, 0.95, stream))
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        reps = 20
        big = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if T * N * 20 < (200 << 20) else None
        times = []
        for _ in range(reps):
            if big is not None:
                big.zero_()  # flush L2 for the sizes that would otherwise be served from it
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            call()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = sorted(times)[reps // 2]
        nbytes = 20 * T * N + 4 * N
        rows.append({"T": T, "N": N, "ms": ms, "algorithmic_mb": nbytes / 1e6, "achieved_gbs": nbytes / ms / 1e6,
                     "frac_of_hbm_peak": nbytes / ms / 1e6 / peak})
        if big is not None:
            # small rollouts: an event pair around ONE launch carries several microseconds of its own; 32 calls replayed
            # from a CUDA graph give the kernel's own time (inputs L2-resident, as they are right after a rollout)
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                sptr = side.cuda_stream
                for _ in range(2):
                    _lib.check(lib.merlin_gae(rew.data_ptr(), val.data_ptr(), done.data_ptr(), last.data_ptr(),
                                              adv.data_ptr(), ret.data_ptr(), T, N, 0.99, 0.95, sptr))
                side.synchronize()
                with torch.cuda.graph(graph, stream=side):
                    for _ in range(32):
                        _lib.check(lib.merlin_gae(rew.data_ptr(), val.data_ptr(), done.data_ptr(), last.data_ptr(),
                                                  adv.data_ptr(), ret.data_ptr(), T, N, 0.99, 0.95, sptr))
            graph.replay()
            torch.cuda.synchronize()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(4):
                graph.replay()
            g1.record()
            torch.cuda.synchronize()
            rows[-1]["us_per_call_cuda_graph_l2_warm"] = g0.elapsed_time(g1) / 128 * 1e3
        print(json.dumps(rows[-1]), flush=True)
        del rew, val, done, last
    # the reference loop (0-dim torch tensors, CPU) on one env, T = 2048
    from oracle import merlin_ref as mr
    T = 2048
    r, v, d = torch.rand(T), torch.randn(T), (torch.rand(T) < 0.01).float()
    t0 = time.perf_counter()
    mr.gae_ppo(r, v, d, 0.3, 0.99, 0.95)
    cpu_ms = (time.perf_counter() - t0) * 1e3
    out = {"what": "GAE + returns kernel", "hbm_peak_gbs": peak, "rows": rows,
           "cpu_reference_loop": {"T": T, "N": 1, "ms": cpu_ms, "kind": "port of src/ppo.py:107-120 (torch CPU)"}}
    print(json.dumps(out["cpu_reference_loop"]))
    if a.out:
        with open(a.out, "w") as f:
            json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
