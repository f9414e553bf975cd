#!/usr/bin/env python
"""Multi-GPU check of the NVLink gradient exchange (run under torchrun on N GPUs of one box).

1. tp_p2p_allreduce_mask (one-shot, two-shot) and tp_p2p_allreduce_nvls (when the symmetric allocation has a multicast
   address) with a mask: against the oracle's fixed rank-order mean (bit-exact; NVLS: bit-exact for W = 2, within
   4 eps of the mean summand magnitude for W > 2 because the switch fixes the summation order) and replica bit-identity.
2. The overlapped reducer inside a real backward pass (ResNet-18, per-rank data): gradients bit-identical to the
   non-overlapped launch order; rank-0 mask broadcast after a rank-dependent pruning step.
3. Timing of a ResNet-50-sized bucket (102 MB) per algorithm against NCCL all_reduce (skipped with --no-timing).
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
from turboprune_b200.grad_exchange import P2PGradReducer
from oracle.train import allreduce_mean_mask


def check_kernels(rank, world, dev):
    ok = True
    probe = P2PGradReducer([torch.nn.Parameter(torch.zeros(1 << 20, device=dev))], algo="two_shot")
    has_mc = probe._bk[0]["mc"] != 0
    del probe
    algos = ["one_shot", "two_shot"] + (["nvls"] if has_mc else [])
    if rank == 0:
        print(f"world={world} multicast={'yes' if has_mc else 'no'} algos={algos}", flush=True)
    for algo in algos:
        for sizes in ([1000, 37, 4096 * 5 + 3], [2_000_000, 513, 7_000_001]):
            params = [torch.nn.Parameter(torch.zeros(s, device=dev)) for s in sizes]
            masks = {id(params[0]): (torch.rand(sizes[0], generator=torch.Generator().manual_seed(9)) < 0.5).float().to(dev)}
            red = P2PGradReducer(params, bucket_cap_mb=16.0, algo=algo, masks=masks, overlap=False)
            for it in range(3):
                gens = [torch.Generator().manual_seed(100 * r + it) for r in range(world)]
                all_g = [[torch.randn(s, generator=gens[r]) for s in sizes] for r in range(world)]
                for p, g in zip(params, all_g[rank]):
                    p.grad = g.to(dev)
                red.reduce()
                torch.cuda.synchronize()
                red.check_status()
                for i, p in enumerate(params):
                    ref = allreduce_mean_mask([all_g[r][i].numpy() for r in range(world)],
                                              masks[id(p)].cpu().numpy() if id(p) in masks else None)
                    got = p.grad.cpu().numpy()
                    if algo == "nvls" and world > 2:
                        # the switch fixes the order of the W-term sum: rounding differs by a few ulp of the partial sums
                        mag = sum(np.abs(all_g[r][i].numpy()) for r in range(world)) / world
                        good = bool(np.all(np.abs(got - ref) <= 4 * np.finfo(np.float32).eps * mag + 1e-30))
                    else:
                        good = np.array_equal(got, ref)
                    if not good:
                        ok = False
                        print(f"[rank {rank}] MISMATCH algo={algo} sizes={sizes} it={it} param={i} maxerr={np.abs(got-ref).max()}", flush=True)
            flat = torch.cat([p.grad.reshape(-1) for p in params])             # replica bit-identity
            gathered = [torch.empty_like(flat) for _ in range(world)]
            dist.all_gather(gathered, flat)
            same = all(torch.equal(gathered[0], g) for g in gathered)
            if not same:
                print(f"[rank {rank}] REPLICAS DIFFER algo={algo} sizes={sizes}", flush=True)
            ok &= same
            red.close(); del red
    return ok, algos


def check_overlap_and_mask_sync(rank, world, dev):
    """The reducer armed inside a real backward pass == the same kernels launched after it; masks imposed from rank 0."""
    import refshim
    from turboprune_b200 import ops
    from turboprune_b200.utils import custom_models as cm, pruning_utils as pu
    ok = True
    torch.manual_seed(0)
    model = cm.TorchVisionModel(refshim.make_cfg("resnet18", "imagenet", precision="bfloat16"))
    torch.manual_seed(1)
    pu.prune_er_erk(model, 0.3)
    model = model.to(dev).train()
    g = torch.Generator().manual_seed(50 + rank)                                 # per-rank data
    x = torch.randn(8, 3, 64, 64, generator=g).to(dev); t = torch.randint(0, 1000, (8,), generator=g).to(dev)
    red = P2PGradReducer(list(model.parameters()), bucket_cap_mb=4.0, algo="two_shot")
    red.set_model_masks(model.model)
    stager = ops.WeightStager([m for _, m in model._masked()])
    res = []
    # (overlap, weight gradients on the side stream): the last two are the harness's configuration — a bucket then holds
    # gradients written on two streams, and its kernel has to wait for both
    for overlap, side in ((False, False), (True, False), (True, False), (True, True), (True, True), (True, True)):
        red.overlap = overlap
        red.zero(); red.arm() if overlap else None
        stager.stage()
        ops.set_wgrad_side_stream(side)
        try:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                torch.nn.functional.cross_entropy(model(x), t).backward()
        finally:
            ops.set_wgrad_side_stream(False)
        launched = sum(red._launched) if overlap else 0
        red.reduce()
        torch.cuda.synchronize(); red.check_status()
        res.append(torch.cat([p.grad.reshape(-1).clone() for p in model.parameters()]))
        if overlap and rank == 0:
            print(f"overlap: {launched} of {len(red._bk)} buckets were launched during the backward pass", flush=True)
        if overlap and launched == 0:
            ok = False
    ok &= all(torch.equal(res[0], r) for r in res[1:])
    gathered = [torch.empty_like(res[1]) for _ in range(world)]
    dist.all_gather(gathered, res[1])
    ok &= all(torch.equal(gathered[0], q) for q in gathered)
    for _, m in model._masked():                                                 # masked weights: exactly zero mean gradient
        ok &= bool((m.weight.grad[m.mask == 0] == 0).all())
    # rank-dependent masks (what SNIP on per-rank batches produces) -> rank 0's masks everywhere
    torch.manual_seed(100 + rank)
    for _, m in model._masked():
        m.mask = (torch.rand_like(m.weight) < 0.5).float()
    pu.sync_masks_from_rank0(model)
    flat = torch.cat([m.mask.reshape(-1) for _, m in model._masked()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    ok &= all(torch.equal(gathered[0], q) for q in gathered)
    red.close()
    return ok


def timing(rank, world, dev, algos):
    n = 25_557_032
    p = [torch.nn.Parameter(torch.zeros(n, device=dev))]
    for algo in algos:
        red = P2PGradReducer(p, bucket_cap_mb=128.0, algo=algo, overlap=False)
        p[0].grad = torch.randn(n, device=dev)
        for _ in range(3):
            red.reduce()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            red.reduce()
        e1.record(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / 10], device=dev); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        ms = float(ms.item())
        if rank == 0:
            print(f"p2p {algo}: {ms*1e3:.1f} us per 102 MB bucket  busbw={2*n*4*(world-1)/world/ms/1e6:.1f} GB/s", flush=True)
        red.close(); del red
    t = torch.randn(n, device=dev)
    for _ in range(3):
        dist.all_reduce(t)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        dist.all_reduce(t)
    e1.record(); torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 10], device=dev); dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        ms = float(ms.item())
        print(f"nccl all_reduce: {ms*1e3:.1f} us  busbw={2*n*4*(world-1)/world/ms/1e6:.1f} GB/s", flush=True)


def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    ok, algos = check_kernels(rank, world, dev)
    ok &= check_overlap_and_mask_sync(rank, world, dev)
    if "--no-timing" not in sys.argv:
        timing(rank, world, dev, algos)
    okt = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("P2P CHECK", "PASS" if int(okt.item()) == 1 else "FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(okt.item()) == 1 else 1)


if __name__ == "__main__":
    main()
