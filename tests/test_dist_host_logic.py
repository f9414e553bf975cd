"""CPU, world_size = 2 over gloo: the host-side logic of the N>1 path (bucket plan identical on all ranks, shard
bounds cover the bucket exactly once, fixed-order mean == the reference DDP's divide-then-sum for W = 2)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from turboprune_b200.grad_exchange import plan_buckets, shard_bounds
        from oracle.train import allreduce_mean_mask
        numels = [1000, 37, 4096 * 5 + 3, 2_000_001, 64, 9408]
        plan = plan_buckets(numels, 1 << 20)
        plans = [None] * world
        dist.all_gather_object(plans, plan)
        assert all(p == plans[0] for p in plans)
        for idx, offs, total in plan:
            assert total % 4 == 0 and all(o % 4 == 0 for o in offs)
            sb = shard_bounds(total, world)
            assert sb[0][0] == 0 and sb[-1][1] == total // 4 * 4
            assert all(sb[i][1] == sb[i + 1][0] for i in range(world - 1))
        # fixed-order mean vs the reference's DDP arithmetic (bucket = grad / W, then SUM) for W = 2: bit-identical
        g = torch.Generator().manual_seed(rank)
        mine = torch.randn(5001, generator=g)
        gathered = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)
        mask = (torch.rand(5001, generator=torch.Generator().manual_seed(7)) < 0.5).float()
        ours = allreduce_mean_mask([t.numpy() for t in gathered], mask.numpy())
        ddp = mine / world
        dist.all_reduce(ddp, op=dist.ReduceOp.SUM)
        assert np.array_equal(ours, (ddp * mask).numpy())
        # pruning on every rank + rank-0 broadcast: rank-dependent masks (what SNIP on per-rank batches gives) end up
        # identical everywhere and equal to rank 0's (the reference prunes on rank 0 and lets DDP broadcast)
        import torch.nn as nn
        from turboprune_b200.utils import mask_layers as ml, pruning_utils as pu
        torch.manual_seed(0)
        net = nn.Sequential(ml.ConvMask(in_channels=8, out_channels=16, kernel_size=3), ml.Conv1dMask(16, 10, bias=True))
        torch.manual_seed(10 + rank)
        for m in net:
            m.mask = (torch.rand_like(m.weight) < 0.5).float()
        mine0 = [m.mask.clone() for m in net]
        pu.sync_masks_from_rank0(net)
        flat = torch.cat([m.mask.reshape(-1) for m in net])
        got = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(got, flat)
        assert all(torch.equal(got[0], t) for t in got)
        if rank == 0:
            assert all(torch.equal(a, m.mask) for a, m in zip(mine0, net))
        assert all(m.mask.shape == m.weight.shape and m.mask.dtype == torch.float32 for m in net)
        q.put((rank, "ok"))
    except Exception as e:      # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world2_gloo_host_logic():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
