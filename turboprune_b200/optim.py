"""Fused SGD(momentum, weight_decay) — one kernel launch for all parameters.

Same update rule and state layout as ``torch.optim.SGD`` as the reference configures it
(harness_definitions/standard_pruning_harness.py:70-75): ``g += wd*w; buf = mu*buf + g
(buf = g on the first step); w -= lr*buf``.  Masked weights keep decaying because the decay
acts on ``w`` itself.  ``state[p]['momentum_buffer']`` is kept so ``state_dict()`` stays
interchangeable with torch's optimizer (the reference saves optimizer_init.pt / _rewind.pt).
"""
import torch

from . import ops


class FusedSGD(torch.optim.Optimizer):
    def __init__(self, params, lr=0.1, momentum=0.0, weight_decay=0.0):
        defaults = dict(lr=lr, momentum=momentum, weight_decay=weight_decay, dampening=0, nesterov=False,
                        maximize=False, foreach=None, differentiable=False, fused=None)
        super().__init__(params, defaults)
        self._lr_dev = {}

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
            first = any("momentum_buffer" not in self.state[p] for p in ps)
            if first:
                if not all("momentum_buffer" not in self.state[p] for p in ps):
                    # mixed (a parameter got its first gradient late): step the newcomers separately
                    old = [p for p in ps if "momentum_buffer" in self.state[p]]
                    new = [p for p in ps if "momentum_buffer" not in self.state[p]]
                    self._launch(gi, group, old, False)
                    self._launch(gi, group, new, True)
                    continue
            self._launch(gi, group, ps, first)
        return loss

    def _launch(self, gi, group, ps, first):
        dev = ps[0].device
        lr_dev = self._lr_dev.get((gi, dev))
        if lr_dev is None:
            lr_dev = self._lr_dev[(gi, dev)] = torch.empty((), dtype=torch.float32, device=dev)
        lr_dev.fill_(float(group["lr"]))
        for p in ps:
            if "momentum_buffer" not in self.state[p]:
                self.state[p]["momentum_buffer"] = torch.empty_like(p, memory_format=torch.contiguous_format)
        grads = [p.grad if p.grad.is_contiguous() else p.grad.contiguous() for p in ps]
        bufs = [self.state[p]["momentum_buffer"] for p in ps]
        ops.sgd_momentum_step(ps, grads, bufs, lr_dev, group["momentum"], group["weight_decay"], first)
