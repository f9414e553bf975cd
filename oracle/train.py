"""Oracle: optimizer / gradient-exchange / train-step restatements (test infrastructure only).

  sgd_momentum_step    torch.optim.SGD as configured at
                       harness_definitions/standard_pruning_harness.py:70-75
                       (g += wd*w; buf = mu*buf + g [buf = g on the first step]; w -= lr*buf).
                       Masked weights keep decaying: wd acts on w, not on mask*w.
  allreduce_mean_mask  DDP gradient mean (harness_definitions/base_harness.py:81 ->
                       c10d Reducer: bucket = grad / W, allreduce SUM).  We sum in fixed
                       rank order and scale by 1/W once; for W a power of two the two are
                       bit-identical, otherwise equal to 1 ulp (tolerance stated in tests).
  train_step           harness_definitions/base_harness.py:115-134
  triangular_schedule  utils/schedulers.py:79-117 (np.interp over [0, warmup, total] -> [0.2, 1, 0], LambdaLR)
  train_epoch          harness_definitions/base_harness.py:151-202 (per-iteration scheduler.step() for the
                       OneCycleLR / Triangular / Trapezoidal schedules, epoch loss = mean of the per-step losses,
                       accuracy over all samples of the epoch in percent)
"""
import numpy as np
import torch


def sgd_momentum_step(w, g, buf, lr, momentum, weight_decay, first_step):
    """One SGD step on numpy fp32 arrays; returns (w_new, buf_new)."""
    w = np.asarray(w, np.float32)
    g = np.asarray(g, np.float32)
    lr32, mu32, wd32 = np.float32(lr), np.float32(momentum), np.float32(weight_decay)
    d = g + wd32 * w if weight_decay != 0 else g.copy()
    if first_step or buf is None:
        nbuf = d.copy()
    else:
        nbuf = mu32 * np.asarray(buf, np.float32) + d
    return (w - lr32 * nbuf).astype(np.float32), nbuf.astype(np.float32)


def allreduce_mean_mask(per_rank_grads, mask=None):
    """Fixed-order sum over ranks, times 1/W, times mask (fp32)."""
    acc = np.asarray(per_rank_grads[0], np.float32).copy()
    for g in per_rank_grads[1:]:
        acc = acc + np.asarray(g, np.float32)
    acc = acc * np.float32(1.0 / len(per_rank_grads))
    if mask is not None:
        acc = acc * np.asarray(mask, np.float32)
    return acc.astype(np.float32)


def train_step(model, optimizer, inputs, targets, amp_dtype=torch.bfloat16, use_amp=True,
               device_type="cpu"):
    """zero_grad -> autocast fwd -> CE -> backward -> step; returns the loss as a float.

    base_harness.py:115-134 minus logging (wandb / torchmetrics are out of scope).
    """
    optimizer.zero_grad()
    with torch.autocast(device_type=device_type, dtype=amp_dtype, enabled=use_amp):
        out = model(inputs)
        loss = torch.nn.functional.cross_entropy(out, targets)
    loss.backward()
    optimizer.step()
    return float(loss.item()), out.detach()


def triangular_schedule(optimizer, steps_per_epoch, epochs, warmup_fraction):
    """utils/schedulers.py:79-117: LambdaLR over the interpolated table (index = number of scheduler steps taken)."""
    total = epochs * steps_per_epoch
    table = np.interp(np.arange(1 + total), [0, int(warmup_fraction * total), total], [0.2, 1, 0])
    return torch.optim.lr_scheduler.LambdaLR(optimizer, table.__getitem__)


def train_epoch(model, optimizer, scheduler, loader, amp_dtype=torch.bfloat16, use_amp=True, device_type="cpu"):
    """base_harness.py:151-202 for a per-iteration schedule: returns (mean loss, accuracy %, per-step lrs, losses)."""
    model.train()
    total, correct, seen, lrs, losses = 0.0, 0, 0, [], []
    for x, t in loader:
        lrs.append(optimizer.param_groups[0]["lr"])
        loss, out = train_step(model, optimizer, x, t, amp_dtype, use_amp, device_type)
        total += loss; losses.append(loss)
        correct += int((out.argmax(1) == t).sum()); seen += int(t.numel())
        scheduler.step()
    return total / len(loader), 100.0 * correct / max(seen, 1), lrs, losses
