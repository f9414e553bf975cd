"""Generic train / eval loop — drop-in for the reference's ``harness_definitions/base_harness.py``.

Same attributes and methods the driver relies on (``.model .train_loader .val_loader .distributed .console
.optimizer .scheduler``; ``train_step / test_step / train_epoch / test``, reference :115-245).  B200 differences:
  * the model is NOT wrapped in DistributedDataParallel: gradients are averaged by ``P2PGradReducer`` (one NVLink
    kernel per bucket) right after ``loss.backward()``; masks are never broadcast (they are replica-identical);
  * the per-step ``loss.item()`` host sync (reference :134) is deferred: losses accumulate on the device and are
    read once per epoch;
  * accuracy is a two-integer device counter instead of torchmetrics (not installed).
"""
from contextlib import nullcontext

import torch
import torch.distributed as dist
import torch.nn as nn
from torch.amp import autocast


class _Console:
    def print(self, *a, **k):
        print(*[str(x) for x in a])

    def rule(self, title=""):
        print("-" * 20, title, "-" * 20)


class _Accuracy:
    def __init__(self, device):
        self.stat = torch.zeros(2, dtype=torch.int64, device=device)

    def update(self, outputs, targets):
        self.stat[0] += (outputs.argmax(1) == targets).sum()
        self.stat[1] += targets.numel()

    def compute(self, distributed=False):
        s = self.stat.clone()
        if distributed:
            dist.all_reduce(s)
        return (s[0].float() / s[1].clamp(min=1).float())

    def reset(self):
        self.stat.zero_()


class BaseHarness:
    def __init__(self, cfg, device, model=None, distributed: bool = False):
        self.cfg = cfg
        self.device = device
        self.distributed = distributed
        self.epoch_counter = 0
        self.console = _Console()
        self.model = self._setup_model(model)
        self.criterion = nn.CrossEntropyLoss()
        self.train_accuracy = _Accuracy(self.device)
        self.test_accuracy = _Accuracy(self.device)
        self.train_loader, self.val_loader = self._setup_dataloaders()
        self.precision, self.use_amp = self._get_dtype_amp()
        self.reducer = None
        self._arena = None
        self._stager = None

    # the reference wraps in DDP here (base_harness.py:74-82); we keep the bare module and reduce explicitly
    def _setup_model(self, model):
        if model is None:
            model = self._create_model()
        return model.to(self.device)

    def _create_model(self):
        raise NotImplementedError

    def _setup_dataloaders(self):
        raise NotImplementedError

    def _get_dtype_amp(self):
        table = {"bfloat16": (torch.bfloat16, True), "float16": (torch.float16, True), "float32": (torch.float32, False)}
        return table.get(self.cfg.experiment_params.training_precision, (torch.float32, False))

    def _ensure_reducer(self):
        if self.distributed and self.reducer is None:
            from ..grad_exchange import P2PGradReducer
            self.reducer = P2PGradReducer(list(self.model.parameters()))

    def _grad_store(self):
        if self.distributed:
            self._ensure_reducer()
            return self.reducer
        if self._arena is None:
            from ..grad_exchange import GradArena
            self._arena = GradArena(list(self.model.parameters()))
        return self._arena

    def _weight_stager(self):
        if self._stager is None:
            from .. import ops
            from ..utils.mask_layers import MASKED_LAYER_TYPES
            self._stager = ops.WeightStager([m for m in self.model.modules() if isinstance(m, MASKED_LAYER_TYPES)])
        return self._stager

    def train_step(self, batch):
        """zero_grad -> autocast forward -> CE -> backward (+ P2P gradient mean) -> SGD (reference :115-134).
        Returns the loss as a 0-dim device tensor (no host sync)."""
        inputs, targets = batch
        inputs, targets = inputs.to(self.device, non_blocking=True), targets.to(self.device, non_blocking=True)
        if inputs.is_cuda:
            # zero_grad: one memset of the persistent gradient storage (param.grad views its slot, the masked layers
            # and fused BN write their gradients straight into it); bf16 weight shadow of all layers: one launch
            self._grad_store().zero()
            self._weight_stager().stage()
        else:
            self.optimizer.zero_grad(set_to_none=True)
        with autocast(device_type="cuda", dtype=self.precision, enabled=self.use_amp):
            outputs = self.model(inputs)
            loss = self.criterion(outputs, targets)
        loss.backward()
        if self.distributed:
            self._ensure_reducer()
            self.reducer.reduce()
        self.optimizer.step()
        self.train_accuracy.update(outputs.detach(), targets)
        return {"loss": loss.detach()}

    def test_step(self, batch):
        inputs, targets = batch
        inputs, targets = inputs.to(self.device, non_blocking=True), targets.to(self.device, non_blocking=True)
        with torch.no_grad(), autocast(device_type="cuda", dtype=self.precision, enabled=self.use_amp):
            outputs = self.model(inputs)
            loss = self.criterion(outputs, targets)
            self.test_accuracy.update(outputs, targets)
        return {"loss": loss.detach()}

    def _epoch(self, loader, step_fn, acc, per_iter_sched):
        total = torch.zeros((), device=self.device)
        for batch in loader:
            total += step_fn(batch)["loss"].float()
            if per_iter_sched and self.scheduler is not None:
                self.scheduler.step()
        avg = total / max(1, len(loader))
        accuracy = acc.compute(self.distributed)
        if self.distributed:
            dist.all_reduce(avg, op=dist.ReduceOp.AVG)
        acc.reset()
        return float(avg.item()), float(accuracy.item()) * 100          # the only host syncs of the epoch

    def train_epoch(self):
        self.model.train()
        per_iter = self.cfg.optimizer_params.scheduler_type in ("OneCycleLR", "TriangularSchedule", "TrapezoidalSchedule")
        loss, acc = self._epoch(self.train_loader, self.train_step, self.train_accuracy, per_iter)
        return {"train_loss": loss, "train_acc": acc}

    def test(self):
        self.model.eval()
        if self.distributed:        # the reference's DDP broadcasts rank 0's BN statistics on every forward
            for m in self.model.modules():
                if isinstance(m, nn.modules.batchnorm._BatchNorm) and m.running_mean is not None:
                    dist.broadcast(m.running_mean, 0); dist.broadcast(m.running_var, 0)
        loss, acc = self._epoch(self.val_loader, self.test_step, self.test_accuracy, False)
        return {"test_loss": loss, "test_acc": acc}
