"""Quick device timing of the fused step kernel (CUDA events) over batch sizes / output modes."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "ppo-2dgrid_b200"))
import numpy as np, torch
from merlin_b200 import BatchedMerlinEnv, layouts

def main():
    Ns = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else "4096,65536,262144,1048576".split(","))]
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100
    L = 4096
    t0 = time.time()
    cells, agent = layouts.generate("mediumhard", 16, range(777_000_000, 777_000_000 + L))
    print(f"layouts: {L} in {time.time()-t0:.1f}s", flush=True)
    dev = "cuda:0"
    for N in Ns:
        for mode in ("rgb", "rgb+sym", "sym"):
            env = BatchedMerlinEnv(N, cells, agent, width=16, height=16, device=dev,
                                   want_rgb="rgb" in mode, want_symbolic="sym" in mode)
            env.reset()
            acts = torch.randint(0, 3, (16, N), device=dev)
            for i in range(10):
                env.step(acts[i % 16])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                env.step(acts[i % 16])
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            sps = N / (ms * 1e-3)
            bytes_per = (9408 if "rgb" in mode else 0) + (147 if "sym" in mode else 0) + 256 + 32 + 8 + 6
            print(f"N={N:8d} {mode:8s} {ms*1e3:9.1f} us/step  {sps:.3e} steps/s  {sps*bytes_per/1e9:8.1f} GB/s", flush=True)
            env.close()

if __name__ == "__main__":
    main()
