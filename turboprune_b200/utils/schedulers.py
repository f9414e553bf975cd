"""LR schedule used by every shipped config of the reference: TriangularSchedule (utils/schedulers.py:79-117)."""
import numpy as np
import torch


def TriangularSchedule(cfg, optimizer, steps_per_epoch, epochs_per_level=None):
    epochs = epochs_per_level if epochs_per_level is not None else cfg.experiment_params.epochs_per_level
    total = epochs * steps_per_epoch
    table = np.interp(np.arange(1 + total), [0, int(cfg.optimizer_params.warmup_fraction * total), total], [0.2, 1, 0])
    return torch.optim.lr_scheduler.LambdaLR(optimizer, lambda i: float(table[min(i, total)]))
