"""Fused SGD(momentum, weight_decay) — one kernel launch for all parameters.

Same update rule and state layout as ``torch.optim.SGD`` as the reference configures it
(harness_definitions/standard_pruning_harness.py:70-75): ``g += wd*w; buf = mu*buf + g
(buf = g on the first step); w -= lr*buf``.  Masked weights keep decaying because the decay
acts on ``w`` itself.  ``state[p]['momentum_buffer']`` is kept so ``state_dict()`` stays
interchangeable with torch's optimizer (the reference saves optimizer_init.pt / _rewind.pt).

CUDA-graph friendly: the learning rate lives in a device scalar (``sync_lr()`` copies
``param_groups[i]['lr']`` into it; inside a captured step nothing is baked in), and the device-side
pointer table is re-uploaded only when a parameter / gradient / buffer pointer changed.
"""
import torch

from . import _cabi, ops


class FusedSGD(torch.optim.Optimizer):
    def __init__(self, params, lr=0.1, momentum=0.0, weight_decay=0.0, capturable=False):
        defaults = dict(lr=lr, momentum=momentum, weight_decay=weight_decay, dampening=0, nesterov=False,
                        maximize=False, foreach=None, differentiable=False, fused=None)
        super().__init__(params, defaults)
        self.capturable = capturable
        self._lr_dev = {}
        self._table = {}          # (group, first) -> (pointer signature, workspace tensor)

    def _lr_tensor(self, gi, dev):
        t = self._lr_dev.get((gi, dev))
        if t is None:
            t = self._lr_dev[(gi, dev)] = torch.full((), float(self.param_groups[gi]["lr"]), dtype=torch.float32, device=dev)
        return t

    def sync_lr(self):
        """Copy every group's host lr into its device scalar (call between graph replays)."""
        for (gi, dev), t in self._lr_dev.items():
            t.fill_(float(self.param_groups[gi]["lr"]))

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        for gi, group in enumerate(self.param_groups):
            ps = [p for p in group["params"] if p.grad is not None]
            if not ps:
                continue
            new = [p for p in ps if "momentum_buffer" not in self.state[p]]
            if new and len(new) != len(ps):
                # a parameter got its first gradient late: step the newcomers separately
                self._launch(gi, group, [p for p in ps if "momentum_buffer" in self.state[p]], False)
                self._launch(gi, group, new, True)
            else:
                self._launch(gi, group, ps, bool(new))
        return loss

    def _launch(self, gi, group, ps, first):
        dev = ps[0].device
        lr_dev = self._lr_tensor(gi, dev)
        if not self.capturable:
            lr_dev.fill_(float(group["lr"]))
        for p in ps:
            if "momentum_buffer" not in self.state[p]:
                self.state[p]["momentum_buffer"] = torch.empty_like(p, memory_format=torch.contiguous_format)
        grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in ps]
        bufs = [self.state[p]["momentum_buffer"] for p in ps]
        sig = tuple(t.data_ptr() for t in ps) + tuple(t.data_ptr() for t in grads) + tuple(t.data_ptr() for t in bufs)
        key = (gi, dev)
        cached = self._table.get(key)
        if cached is None or cached[1].numel() < _cabi.load().tp_segtable_workspace_bytes(len(ps)):
            ws = torch.empty(_cabi.load().tp_segtable_workspace_bytes(len(ps)), dtype=torch.uint8, device=dev)
            cached = (None, ws)
        hit = cached[0] == sig
        ops.sgd_momentum_step(ps, grads, bufs, lr_dev, group["momentum"], group["weight_decay"], first,
                              table_ws=cached[1], table_cached=hit)
        self._table[key] = (sig, cached[1])
