"""Oracle: reference-equivalent masked model on CPU (test infrastructure only).

Builds the same module graph the reference builds in utils/custom_models.py
(TorchVisionModel :169-220): a torchvision network, CIFAR stem/classifier surgery
(:197-215), then every nn.Conv2d replaced by a conv that multiplies its weight by an
fp32 ``mask`` buffer (utils/mask_layers.py:19-34) and every nn.Linear by the k=1
conv1d form with weight [out, in, 1] (utils/mask_layers.py:94-119).  State-dict keys
and shapes match the reference (``<layer>.weight``, ``<layer>.mask``), so checkpoints
move freely between oracle, reference and product.

Used by tests / smoke (checker) and by bench.py's cpu_baseline / --impl reference leg
(the timed CPU implementation of the path) — never by the product.
"""
import torch
import torch.nn as nn
from torchvision import models as tvm

from . import mask_ops


class RefMaskedConv2d(nn.Conv2d):
    def __init__(self, **kw):
        super().__init__(**kw)
        self.register_buffer("mask", torch.ones_like(self.weight))

    def forward(self, x):
        w = self.mask.to(self.weight.device) * self.weight
        return nn.functional.conv2d(x, w, self.bias, self.stride, self.padding, self.dilation, self.groups)


class RefMaskedFC(nn.Conv1d):
    """nn.Linear stand-in with weight [out, in, 1] (reference Conv1dMask)."""

    def __init__(self, in_features, out_features, bias=False):
        super().__init__(in_features, out_features, kernel_size=1, stride=1, bias=bias)
        self.register_buffer("mask", torch.ones_like(self.weight))

    def forward(self, x):
        w = self.mask.to(self.weight.device) * self.weight
        return nn.functional.conv1d(x.unsqueeze(-1), w, self.bias).squeeze(-1)


MASKED_TYPES = (RefMaskedConv2d, RefMaskedFC)


def _swap(module):
    for name, child in module.named_children():
        if isinstance(child, nn.Linear):
            setattr(module, name, RefMaskedFC(child.in_features, child.out_features, child.bias is not None))
        elif isinstance(child, nn.Conv2d):
            # like the reference (custom_models.py:86-94) groups/dilation are not forwarded
            setattr(module, name, RefMaskedConv2d(
                in_channels=child.in_channels, out_channels=child.out_channels,
                kernel_size=child.kernel_size, stride=child.stride, padding=child.padding,
                bias=child.bias is not None))
        else:
            _swap(child)


def build(model_name="resnet50", dataset="imagenet"):
    net = getattr(tvm, model_name)(weights=None)
    ds = dataset.lower()
    if ds in ("cifar10", "cifar100"):
        ncls = 10 if ds == "cifar10" else 100
        if model_name.startswith("resnet"):
            net.conv1 = nn.Conv2d(3, 64, kernel_size=3, stride=1, padding=1, bias=False)
            net.maxpool = nn.Identity()
            net.fc = nn.Linear(net.fc.in_features, ncls)
        elif model_name.startswith("vgg"):
            net.features[0] = nn.Conv2d(3, 64, kernel_size=3, padding=1)
            net.classifier[-1] = nn.Linear(net.classifier[-1].in_features, ncls)
    _swap(net)
    return net


def masked_layers(net):
    return [(n, m) for n, m in net.named_modules() if isinstance(m, MASKED_TYPES)]


def set_er_masks(net, probs):
    """set_er_mask for every layer — utils/mask_layers.py:36-43 (torch Philox/MT stream)."""
    for (_, m), p in zip(masked_layers(net), probs):
        m.mask = torch.zeros_like(m.weight).bernoulli_(p)
