"""Synthetic on-device generator standing in for the reference's loaders (utils/dataset.py: CifarLoader
:101-226, FFCVImagenet :347-430 — out of scope: they need the data sets / FFCV / network).

Same batch contract: an iterable of ``(images fp32 [B,3,H,W], labels int64 [B])`` with ``len()``; ImageNet-shaped
batches come channels_last like FFCV's ToTorchImage.  Seeded per rank; a fixed number of distinct batches is
generated once and cycled so an epoch costs no host work.
"""
import torch


class SyntheticLoader:
    def __init__(self, batch_size, steps, shape, num_classes, device, seed=0, distinct=4, channels_last=False):
        g = torch.Generator(device=device).manual_seed(seed)
        self.steps = steps
        self.batches = []
        for _ in range(min(distinct, steps)):
            x = torch.randn(batch_size, *shape, device=device, generator=g)
            if channels_last:
                x = x.contiguous(memory_format=torch.channels_last)
            self.batches.append((x, torch.randint(0, num_classes, (batch_size,), device=device, generator=g)))

    def __len__(self):
        return self.steps

    def __iter__(self):
        for i in range(self.steps):
            yield self.batches[i % len(self.batches)]


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
        self.train_loader = SyntheticLoader(bs, steps, shape, ncls, device, seed, channels_last=name.startswith("imagenet"))
        self.test_loader = SyntheticLoader(bs, max(1, steps // 4), shape, ncls, device, seed + 7919,
                                           channels_last=name.startswith("imagenet"))
