"""Data-parallel plumbing for the learners: one process per GPU, `torch.distributed` (NCCL over NVLink on the
box, gloo in the CPU tests).  Environments and FOMAML tasks are sharded by rank and never communicate; the
only collectives are (1) the SUM all-reduce of the flattened gradient -- PPO: once per minibatch step before
clipping (reference src/ppo.py:154-156); FOMAML: once per meta-iteration before the division by the task
count (src/fomaml.py:207-209) -- and (2) three scalars for the batch-wide advantage statistics (src/ppo.py:125).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard(items, r=None, w=None):
    """Contiguous, balanced slice of `items` owned by rank r of w (tasks / seeds / env indices)."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    items = list(items)
    base, extra = divmod(len(items), w)
    lo = r * base + min(r, extra)
    return items[lo: lo + base + (1 if r < extra else 0)]


class FlatGrads:
    """All parameters' `.grad` as views into ONE contiguous buffer, so the gradient all-reduce is a single
    collective on memory autograd already wrote (2.98 MB for the CNN actor-critic)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        first = self.params[0]
        self.flat = torch.zeros(n, dtype=first.dtype, device=first.device)
        off = 0
        for p in self.params:
            p.grad = self.flat[off: off + p.numel()].view_as(p)
            off += p.numel()

        self.timing = None  # list of (start, end) CUDA events once enable_timing() was called

    def zero_(self):
        self.flat.zero_()

    def enable_timing(self, on=True):
        """Bracket every gradient all-reduce with CUDA events (measurement only: `collective_seconds()` sums them)."""
        self.timing = [] if on else None

    def collective_seconds(self, clear=True):
        """Device time spent inside the gradient all-reduces since the last call (synchronises)."""
        if not self.timing:
            return 0.0
        torch.cuda.synchronize(self.flat.device)
        total = sum(a.elapsed_time(b) for a, b in self.timing) * 1e-3
        if clear:
            self.timing = []
        return total

    def nbytes(self):
        return self.flat.numel() * self.flat.element_size()

    def all_reduce_mean(self):
        """Average over ranks (each rank's loss is a mean over its own equally sized shard)."""
        w = world_size()
        if w > 1:
            if self.timing is not None:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            self.flat.div_(w)
            if self.timing is not None:
                b.record()
                self.timing.append((a, b))

    def all_reduce_sum(self):
        if world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)


def global_mean_std(x):
    """Mean and unbiased std of `x` over all ranks' elements (equals `x.mean(), x.std()` at world size 1)."""
    w = world_size()
    if w == 1:
        return x.mean(), x.std()
    xd = x.double()
    s = torch.stack([xd.sum(), (xd * xd).sum(), torch.tensor(float(x.numel()), dtype=torch.float64, device=x.device)])
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    n = s[2]
    mean = s[0] / n
    var = (s[1] - n * mean * mean) / (n - 1)
    return mean.to(x.dtype), var.clamp_min(0).sqrt().to(x.dtype)


def broadcast_parameters(module, src=0):
    """Make every rank start from rank `src`'s weights."""
    if world_size() > 1:
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src)
