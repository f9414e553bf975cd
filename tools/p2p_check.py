#!/usr/bin/env python
"""Multi-GPU check of tp_p2p_allreduce_mask (run under torchrun on N GPUs of one box).
Compares with the oracle's fixed-order mean and with NCCL all_reduce; times both."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from ctypes import c_void_p
from turboprune_b200 import _cabi
from turboprune_b200.grad_exchange import P2PGradReducer, PAD_FLOATS
from oracle.train import allreduce_mean_mask

def main():
    rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    ok = True
    for algo in ("one_shot", "two_shot"):
        for sizes in ([1000, 37, 4096 * 5 + 3], [2_000_000, 513, 7_000_001]):
            params = [torch.nn.Parameter(torch.zeros(s, device=dev)) for s in sizes]
            masks = {id(params[0]): (torch.rand(sizes[0], generator=torch.Generator().manual_seed(9)) < 0.5).float().to(dev)}
            red = P2PGradReducer(params, bucket_cap_mb=16.0, algo=algo, masks=masks)
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
                    if not np.array_equal(got, ref):
                        ok = False
                        print(f"[rank {rank}] MISMATCH algo={algo} sizes={sizes} it={it} param={i} maxerr={np.abs(got-ref).max()}")
            # replica bit-identity
            flat = torch.cat([p.grad.reshape(-1) for p in params])
            gathered = [torch.empty_like(flat) for _ in range(world)]
            dist.all_gather(gathered, flat)
            ok &= all(torch.equal(gathered[0], g) for g in gathered)
            del red
    # timing: RN50-sized gradient (102 MB) — P2P kernel vs NCCL all_reduce
    n = 25_557_032
    p = [torch.nn.Parameter(torch.zeros(n, device=dev))]
    for algo in ("one_shot", "two_shot"):
        red = P2PGradReducer(p, bucket_cap_mb=128.0, algo=algo)
        p[0].grad = torch.randn(n, device=dev)
        for _ in range(3):
            red.reduce()
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            red.reduce()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        if rank == 0:
            print(f"p2p {algo}: {ms*1e3:.1f} us per 102 MB bucket  busbw={2*n*4*(world-1)/world/ms/1e6:.1f} GB/s")
        del red
    t = torch.randn(n, device=dev)
    for _ in range(3):
        dist.all_reduce(t)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        dist.all_reduce(t)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    if rank == 0:
        print(f"nccl all_reduce: {ms*1e3:.1f} us  busbw={2*n*4*(world-1)/world/ms/1e6:.1f} GB/s")
        print("P2P CHECK", "PASS" if ok else "FAIL")
    okt = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(okt, op=dist.ReduceOp.MIN)
    dist.destroy_process_group()
    sys.exit(0 if int(okt.item()) == 1 else 1)

if __name__ == "__main__":
    main()
