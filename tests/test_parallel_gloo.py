"""CPU tier, world_size = 2 over gloo: the host-side logic of the multi-GPU path (src/parallel.py) -- rank
sharding of tasks, the single flattened-gradient all-reduce, batch-wide advantage statistics and the weight
broadcast.  On the box the same code runs over NCCL; the env/GAE kernels themselves never communicate."""
from __future__ import annotations

import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from src import parallel
        from src.actor_critic import MLPActorCritic

        assert parallel.world_size() == world and parallel.rank() == rank
        # 1. task sharding: FOMAML's 32 seeds -> 16 per rank, disjoint, ordered
        seeds = list(range(100, 132))
        mine = parallel.shard(seeds)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        assert sum(gathered, []) == seeds and len(mine) == 16

        # 2. weight broadcast: ranks start from different seeds, end identical to rank 0
        torch.manual_seed(rank)
        net = MLPActorCritic(6, 3, hidden_dim=8)
        parallel.broadcast_parameters(net)
        torch.manual_seed(0)
        ref = MLPActorCritic(6, 3, hidden_dim=8)
        for a, b in zip(net.parameters(), ref.parameters()):
            assert torch.equal(a, b)

        # 3. data-parallel gradient == single-process gradient over the concatenated batch
        g = torch.Generator().manual_seed(7)
        x = torch.randn(8, 6, generator=g)
        y = torch.randn(8, generator=g)
        flat = parallel.FlatGrads(net.parameters())
        lo, hi = rank * 4, rank * 4 + 4
        flat.zero_()
        ((net(x[lo:hi])[1] - y[lo:hi]) ** 2).mean().backward()
        flat.all_reduce_mean()
        ((ref(x)[1] - y) ** 2).mean().backward()
        ref_flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in ref.parameters()])
        assert torch.allclose(flat.flat, ref_flat, atol=1e-6), (flat.flat - ref_flat).abs().max()
        for p in net.parameters():  # .grad still aliases the flat buffer after backward + all-reduce
            assert p.grad.data_ptr() >= flat.flat.data_ptr()
        # a second backward accumulates in place (zero_ is what resets it)
        before = flat.flat.clone()
        ((net(x[lo:hi])[1] - y[lo:hi]) ** 2).mean().backward()
        assert not torch.equal(before, flat.flat)
        flat.zero_()
        assert float(flat.flat.abs().sum()) == 0.0

        # 4. FOMAML-style SUM then divide by the global task count
        flat.flat.fill_(float(rank + 1))
        flat.all_reduce_sum()
        assert torch.all(flat.flat == 3.0)

        # 5. batch-wide advantage statistics == statistics of the concatenation
        adv = torch.randn(64, generator=g)
        mean, std = parallel.global_mean_std(adv[rank * 32: rank * 32 + 32])
        assert torch.allclose(mean, adv.mean(), atol=1e-6) and torch.allclose(std, adv.std(), atol=1e-6)
        out.put((rank, "ok"))
    except Exception as e:  # pragma: no cover - surfaced by the parent
        out.put((rank, f"{type(e).__name__}: {e}"))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert results == {0: "ok", 1: "ok"}, results
