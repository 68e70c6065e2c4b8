#!/usr/bin/env python
"""Time the minibatch read path of a symbolic rollout: merlin_env_render (u8 frames, u8 blocked) and
merlin_env_render_f32 (the first layer's float32 input), with a random row gather, against what the learner did before
the float32 kernel existed (u8 blocked render + PyTorch cast).  CUDA events on the launching stream, L2 flushed by the
working set (M = 262 144) or by an explicit 256 MB fill between launches (M = 16 384, the PPO minibatch)."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ppo-2dgrid_b200"))
import torch
from merlin_b200 import BatchedMerlinEnv

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=None)
ap.add_argument("--reps", type=int, default=20)
a = ap.parse_args()
dev = "cuda:0"
peak = 6549.4
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    pass
R = 524288  # stored rows (a 4096-env x 128-step rollout)
env = BatchedMerlinEnv(R, width=16, height=16, device=dev, generate=("mediumhard", 1, 65536), want_symbolic=True)
env.reset()
for _ in range(4):
    env.step(torch.randint(0, 3, (R,), device=dev))
sym = env.obs_symbolic.clone()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, flush_l2):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(a.reps):
        if flush_l2:
            flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


rows = []
for M in (4096, 16384, 262144):
    idx = torch.randperm(R, device=dev)[:M].contiguous()
    out_u8 = torch.empty((M, 56, 56, 3), dtype=torch.uint8, device=dev)
    out_blk = torch.empty((M, 14, 14, 48), dtype=torch.uint8, device=dev)
    out_f32 = torch.empty((M, 14, 14, 48), dtype=torch.float32, device=dev)
    cases = {
        "render u8 frames": (lambda: env.render(sym, idx, out=out_u8), 9408 + 155),
        "render u8 blocked": (lambda: env.render(sym, idx, out=out_blk, blocked=True), 9408 + 155),
        "render_f32 blocked": (lambda: env.render(sym, idx, out=out_f32, blocked=True, dtype=torch.float32), 4 * 9408 + 155),
        "render_f32 blocked /255": (lambda: env.render(sym, idx, out=out_f32, blocked=True, dtype=torch.float32,
                                                      normalise="divide"), 4 * 9408 + 155),
        "u8 blocked + torch .float() (previous learner path)": (
            lambda: env.render(sym, idx, out=out_blk, blocked=True).permute(0, 3, 1, 2).float(), None),
    }
    for name, (fn, bytes_per_frame) in cases.items():
        ms = timed(fn, flush_l2=M * 9408 * 4 < (512 << 20))
        row = {"frames": M, "case": name, "ms": ms}
        if bytes_per_frame:
            row["algorithmic_GBps"] = M * bytes_per_frame / ms / 1e6
            row["frac_of_copy_peak"] = row["algorithmic_GBps"] / peak
        if M <= 16384 and bytes_per_frame:
            # small launches: 16 calls replayed from a CUDA graph give the kernel's own time (no host gaps, L2-warm input)
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.stream(side):
                fn()
                side.synchronize()
                with torch.cuda.graph(graph, stream=side):
                    for _ in range(16):
                        fn()
            graph.replay()
            torch.cuda.synchronize()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(4):
                graph.replay()
            g1.record()
            torch.cuda.synchronize()
            row["us_per_launch_cuda_graph"] = g0.elapsed_time(g1) / 64 * 1e3
            row["frac_of_copy_peak_cuda_graph"] = M * bytes_per_frame / row["us_per_launch_cuda_graph"] / 1e3 / peak
        rows.append(row)
        print(row, flush=True)
res = {"what": "minibatch read path of a symbolic rollout (row gather + render)", "stored_rows": R, "hbm_copy_peak_GBps": peak,
       "timing": "median of %d launches, CUDA events, L2 flushed between launches at M=16384" % a.reps, "rows": rows}
if a.out:
    os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)
