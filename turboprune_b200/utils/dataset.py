"""Synthetic on-device generator standing in for the reference's loaders (utils/dataset.py: CifarLoader
:101-226, FFCVImagenet :347-430 — out of scope: they need the data sets / FFCV / network).

Same batch contract: an iterable of ``(images fp32 [B,3,H,W], labels int64 [B])`` with ``len()``; ImageNet-shaped
batches come channels_last like FFCV's ToTorchImage.  Seeded per rank; either a fixed number of distinct batches is
generated once and cycled (an epoch costs no host work), or (``dataset_params.synthetic_fresh``) every step draws a new
batch on the device.  ``DevicePrefetcher`` is the host->device leg for loaders that produce pinned host batches.
"""
import torch


class SyntheticLoader:
    """``fresh=True``: every iteration draws a new batch on the device (Philox, seeded per rank — SURVEY.md §8(d));
    otherwise ``distinct`` batches are generated once and cycled."""

    def __init__(self, batch_size, steps, shape, num_classes, device, seed=0, distinct=4, channels_last=False, fresh=False):
        self.gen = torch.Generator(device=device).manual_seed(seed)
        self.steps = steps
        self.batch_size, self.shape, self.num_classes, self.device = batch_size, tuple(shape), num_classes, device
        self.channels_last = channels_last
        self.fresh = fresh
        self.batches = [] if fresh else [self._draw() for _ in range(min(distinct, steps))]

    def _draw(self):
        c, h, w = self.shape
        if self.channels_last:       # NHWC memory, logical NCHW (what FFCV's ToTorchImage hands over, dataset.py:391)
            x = torch.randn(self.batch_size, h, w, c, device=self.device, generator=self.gen).permute(0, 3, 1, 2)
        else:
            x = torch.randn(self.batch_size, c, h, w, device=self.device, generator=self.gen)
        return x, torch.randint(0, self.num_classes, (self.batch_size,), device=self.device, generator=self.gen)

    def __len__(self):
        return self.steps

    def __iter__(self):
        for i in range(self.steps):
            yield self._draw() if self.fresh else self.batches[i % len(self.batches)]


class DevicePrefetcher:
    """Wraps an iterable of HOST batches (pinned ``(images, labels)``) and yields device batches: batch i+1 crosses PCIe
    on a copy stream while step i computes (two staging slots, guarded by events).  Stands where the reference's
    loaders hand over device tensors (FFCV ``ToDevice(non_blocking=True)``, dataset.py:385-430)."""

    def __init__(self, host_loader, device):
        self.loader, self.device = host_loader, device
        self.copy_stream = torch.cuda.Stream(device)
        self.slots = [None, None]
        self.ready = [torch.cuda.Event(), torch.cuda.Event()]
        self.consumed = [torch.cuda.Event(), torch.cuda.Event()]

    def __len__(self):
        return len(self.loader)

    def _issue(self, j, batch):
        x, t = batch
        if self.slots[j] is None or self.slots[j][0].shape != x.shape:
            self.slots[j] = (torch.empty_strided(x.shape, x.stride(), dtype=x.dtype, device=self.device),
                             torch.empty(t.shape, dtype=t.dtype, device=self.device))
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.consumed[j])
            self.slots[j][0].copy_(x, non_blocking=True); self.slots[j][1].copy_(t, non_blocking=True)
            self.ready[j].record(self.copy_stream)

    def __iter__(self):
        cur = torch.cuda.current_stream(self.device)
        for ev in self.consumed:
            ev.record(cur)
        it = iter(self.loader)
        nxt = next(it, None)
        if nxt is None:
            return
        self._issue(0, nxt)
        i = 0
        while nxt is not None:
            j = i % 2
            nxt = next(it, None)
            if nxt is not None:
                self._issue(1 - j, nxt)              # overlaps with the step consuming slot j
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(self.ready[j])
            yield self.slots[j]
            self.consumed[j].record(torch.cuda.current_stream(self.device))
            i += 1


class SyntheticLoaders:
    """train_loader / test_loader pair sized from the config (dataset_params.total_batch_size // world_size,
    reference dataset.py:411)."""

    def __init__(self, cfg, device, world_size=1, rank=0):
        name = cfg.dataset_params.dataset_name.lower()
        ncls = 1000 if name.startswith("imagenet") else (100 if name.startswith("cifar100") else 10)
        shape = (3, 224, 224) if name.startswith("imagenet") else (3, 32, 32)
        bs = max(1, cfg.dataset_params.total_batch_size // world_size)
        steps = int(getattr(cfg.dataset_params, "synthetic_steps_per_epoch", 8))
        seed = cfg.experiment_params.seed * world_size + rank
        fresh = bool(getattr(cfg.dataset_params, "synthetic_fresh", False))
        self.train_loader = SyntheticLoader(bs, steps, shape, ncls, device, seed, channels_last=name.startswith("imagenet"),
                                            fresh=fresh)
        self.test_loader = SyntheticLoader(bs, max(1, steps // 4), shape, ncls, device, seed + 7919,
                                           channels_last=name.startswith("imagenet"))
