"""Data-parallel gradient exchange over NVLink peer memory (replaces DDP's Reducer + NCCL).

The reference wraps the model in ``DistributedDataParallel`` with default settings
(harness_definitions/base_harness.py:81): per-bucket ``grad / W`` -> ``ncclAllReduce(SUM)`` ->
copy back, plus a broadcast of every buffer (all masks!) from rank 0 on every forward.  Here:

  * gradients are packed into a few large buckets that live in *symmetric memory*
    (``torch.distributed._symmetric_memory``: every rank's bucket is mapped into every peer's
    address space over NVLink / NVSwitch);
  * one ``tp_p2p_allreduce_mask`` kernel per bucket reads the peers' copies directly, sums them
    in fixed rank order (bit-identical result on every rank), scales by 1/W, applies the mask
    and leaves the result in the local bucket — ``param.grad`` then simply views the bucket;
  * nothing else is exchanged: masks are deterministic functions of replica-identical state, so
    the reference's per-step mask broadcast disappears.  torch.distributed (NCCL) is only used
    for rendezvous and scalar reductions.
"""
import ctypes
from ctypes import c_void_p

import torch
import torch.distributed as dist

from . import _cabi, ops

PAD_FLOATS = 1024          # 4 KiB of signal-pad slots at the head of every symmetric bucket


def plan_buckets(numels, cap_elems):
    """Pure host logic: pack parameters (given in REVERSE registration order, i.e. the order gradients become
    ready) into buckets of at most ``cap_elems`` fp32 slots; every slot is padded to a multiple of 4 elements so it
    stays 16-byte aligned.  Returns [(indices, offsets, total_elems)].  Must be identical on every rank."""
    buckets, cur, offs, total = [], [], [], 0
    for i, n in enumerate(numels):
        slot = (n + 3) // 4 * 4
        if cur and total + slot > cap_elems:
            buckets.append((cur, offs, total)); cur, offs, total = [], [], 0
        cur.append(i); offs.append(total); total += slot
    if cur:
        buckets.append((cur, offs, total))
    return buckets


def shard_bounds(numel, world):
    """float4-granular shard [s0, s1) of every rank for the two-shot algorithm (mirrors k_p2p_allreduce<1>)."""
    n4 = numel // 4
    per = (n4 + world - 1) // world
    return [(min(n4, per * r) * 4, min(n4, min(n4, per * r) + per) * 4) for r in range(world)]


class GradArena:
    """Persistent, flat gradient storage for ONE rank: ``param.grad`` permanently views a 16-byte aligned slot of
    one fp32 buffer, ``zero()`` is a single memset.  Pointers never change, so the fused SGD keeps its device-side
    pointer table and the whole train step can be captured into a CUDA graph (autograd accumulates into the
    zeroed slots: 0 + g == g bit-exactly, same values as the reference's zero_grad + assign)."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        (idx, offs, total), = plan_buckets([p.numel() for p in self.params], 1 << 62)
        self.flat = torch.zeros(total, dtype=torch.float32, device=self.params[0].device)
        self.views = [self.flat[o:o + p.numel()].view_as(p) for o, p in zip(offs, self.params)]
        self.attach()

    def attach(self):
        for p, v in zip(self.params, self.views):
            p.grad = v
            p._tp_grad_slot = v          # masked layers / fused BN write their gradients straight into the slot

    def zero(self):
        self.flat.zero_()
        self.attach()


class P2PGradReducer:
    def __init__(self, params, bucket_cap_mb=64.0, algo="auto", group=None, masks=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group or dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.params = [p for p in params if p.requires_grad]
        self.algo = algo
        dev = self.params[0].device
        self.device = dev
        cap = int(bucket_cap_mb * 1024 * 1024 / 4)
        # reverse order (gradients become ready back to front), like DDP's bucket assignment
        rev = list(reversed(self.params))
        self.plan = plan_buckets([p.numel() for p in rev], cap)
        self.buckets = [[rev[i] for i in idx] for idx, _, _ in self.plan]
        self._bk = []
        masks = masks or {}
        for plist in self.buckets:
            offs, total = [], 0
            for p in plist:
                offs.append(total); total += (p.numel() + 3) // 4 * 4
            buf = symm_mem.empty(PAD_FLOATS + total, dtype=torch.float32, device=dev)
            buf.zero_()
            hdl = symm_mem.rendezvous(buf, self.group)
            ptrs = [int(x) for x in hdl.buffer_ptrs]
            data_ptrs = (c_void_p * self.world)(*[c_void_p(x + PAD_FLOATS * 4) for x in ptrs])
            pad_ptrs = (c_void_p * self.world)(*[c_void_p(x) for x in ptrs])
            data = buf[PAD_FLOATS:]
            views = [data[o:o + p.numel()].view_as(p) for o, p in zip(offs, plist)]
            algo = self._algo_for(total)
            # one-shot: peers read my bucket while I produce the result, so it needs its own
            # output buffer; two-shot finishes in place (only rank r ever reads shard r).
            out = torch.empty(total, dtype=torch.float32, device=dev) if algo == 0 else data
            out_views = [out[o:o + p.numel()].view_as(p) for o, p in zip(offs, plist)]
            mask = None
            if any(id(p) in masks for p in plist):
                mask = torch.ones(total, dtype=torch.float32, device=dev)
                for o, p in zip(offs, plist):
                    if id(p) in masks:
                        mask[o:o + p.numel()] = masks[id(p)].reshape(-1)
            self._bk.append(dict(buf=buf, hdl=hdl, data=data, views=views, numel=total, params=plist,
                                 data_ptrs=data_ptrs, pad_ptrs=pad_ptrs, mask=mask, algo=algo, out=out,
                                 out_views=out_views))
        self.status = torch.zeros(1, dtype=torch.int32, device=dev)
        torch.cuda.synchronize(dev)
        dist.barrier(self.group)

    def _algo_for(self, numel):
        if self.algo == "one_shot":
            return 0
        if self.algo == "two_shot":
            return 1
        return 0 if numel * 4 <= (1 << 20) else 1

    @torch.no_grad()
    def reduce(self):
        """Average gradients across ranks; afterwards every ``param.grad`` views its bucket slot."""
        lib = _cabi.load()
        st = _cabi.stream_ptr(self.device)
        for bk in self._bk:
            grads = [p.grad for p in bk["params"]]
            live = [(v, g) for v, g in zip(bk["views"], grads) if g is not None and g.data_ptr() != v.data_ptr()]
            if live:
                torch._foreach_copy_([v for v, _ in live], [g for _, g in live])
            for v, g in zip(bk["views"], grads):
                if g is None:
                    v.zero_()
            rc = lib.tp_p2p_allreduce_mask(bk["data_ptrs"], bk["pad_ptrs"], self.rank, self.world, bk["numel"],
                                           c_void_p(bk["mask"].data_ptr()) if bk["mask"] is not None else None,
                                           1.0 / self.world, c_void_p(bk["out"].data_ptr()), bk["algo"],
                                           20000, c_void_p(self.status.data_ptr()), st)
            _cabi.check(rc, "tp_p2p_allreduce_mask")
            ops._count()
            for p, v in zip(bk["params"], bk["out_views"]):
                p.grad = v

    def attach(self):
        """Make every ``param.grad`` a persistent view of its bucket slot (stable pointers: CUDA-graph capturable)."""
        for bk in self._bk:
            for p, v in zip(bk["params"], bk["views"]):
                p.grad = v
                p._tp_grad_slot = v

    def zero(self):
        for bk in self._bk:
            bk["data"].zero_()
        self.attach()

    def check_status(self):
        if int(self.status.item()) != 0:
            raise RuntimeError("tp_p2p_allreduce_mask: peer barrier timed out (a rank did not arrive)")
