"""Fused BatchNorm (+ residual) (+ ReLU) for the NHWC bf16 activation path.

``BatchNorm2dB200`` is a drop-in ``nn.BatchNorm2d`` subclass (same parameters / buffers / state-dict
keys); ``fuse_torchvision_blocks`` rebinds the ``forward`` of torchvision's ResNet blocks and VGG
``features`` so that ``bn -> relu`` and ``bn -> (+identity) -> relu`` run as ONE apply kernel (and one
backward pair) instead of separate ATen batch_norm / relu / add kernels.  The module graph the reference
builds (utils/custom_models.py:184) and its state dict are unchanged.
"""
import types
from ctypes import c_void_p

import torch
import torch.nn as nn

from . import _cabi, ops


def _ptr(t):
    return c_void_p(t.data_ptr()) if t is not None else None


class _BNFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, residual, weight, bias, running_mean, running_var, nbt, momentum, eps, training, relu):
        lib = _cabi.load()
        n, c, h, w = x.shape
        xn = ops.to_nhwc_bf16(x, c)
        rn = ops.to_nhwc_bf16(residual, c) if residual is not None else None
        m = n * h * w
        dev = x.device
        z = torch.empty(n, h, w, c, dtype=torch.bfloat16, device=dev)
        save_mean = torch.empty(c, dtype=torch.float32, device=dev) if training else None
        save_invstd = torch.empty(c, dtype=torch.float32, device=dev) if training else None
        wsb = ops._workspace(lib.tp_bn_workspace_bytes(m, c), dev, "bn")
        with torch.cuda.device(dev):
            rc = lib.tp_bn_forward(_ptr(xn), _ptr(rn), _ptr(z), m, c, _ptr(weight), _ptr(bias), _ptr(running_mean),
                                   _ptr(running_var), _ptr(nbt), float(momentum), float(eps), int(training), int(relu),
                                   _ptr(save_mean), _ptr(save_invstd), _ptr(wsb), wsb.numel(), _cabi.stream_ptr(dev))
        _cabi.check(rc, "tp_bn_forward")
        ops._count(3 if training else 2)
        if training:
            ctx.save_for_backward(xn, z if relu else None, weight, save_mean, save_invstd)
            ctx.relu = relu
            ctx.has_res = residual is not None
            ctx.res_dtype = residual.dtype if residual is not None else None
            ctx.x_dtype = x.dtype
        return z.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dz):
        lib = _cabi.load()
        xn, z, weight, save_mean, save_invstd = ctx.saved_tensors
        n, h, w, c = xn.shape
        m = n * h * w
        dev = xn.device
        dzn = ops.to_nhwc_bf16(dz, c)
        dy = torch.empty_like(xn)
        dres = torch.empty_like(xn) if ctx.has_res else None
        dweight = torch.empty(c, dtype=torch.float32, device=dev)
        dbias = torch.empty(c, dtype=torch.float32, device=dev)
        wsb = ops._workspace(lib.tp_bn_workspace_bytes(m, c), dev, "bn")
        with torch.cuda.device(dev):
            rc = lib.tp_bn_backward(_ptr(dzn), _ptr(z), _ptr(xn), m, c, _ptr(weight), _ptr(save_mean), _ptr(save_invstd),
                                    int(ctx.relu), _ptr(dy), _ptr(dres), _ptr(dweight), _ptr(dbias), _ptr(wsb), wsb.numel(),
                                    _cabi.stream_ptr(dev))
        _cabi.check(rc, "tp_bn_backward")
        ops._count(3)
        gx = dy.permute(0, 3, 1, 2)
        if gx.dtype != ctx.x_dtype:
            gx = gx.to(ctx.x_dtype)
        gr = dres.permute(0, 3, 1, 2) if dres is not None else None
        if gr is not None and gr.dtype != ctx.res_dtype:
            gr = gr.to(ctx.res_dtype)
        return gx, gr, dweight, dbias, None, None, None, None, None, None, None


class BatchNorm2dB200(nn.BatchNorm2d):
    """nn.BatchNorm2d whose CUDA path is the fused sm_100a kernel; ``forward(x, residual=None, relu=False)``."""

    def forward(self, x, residual=None, relu=False):
        training = self.training or not self.track_running_stats
        # eval-mode BN inside an autograd graph is not on the hot path: leave it to torch
        if not x.is_cuda or x.shape[1] % 8 != 0 or (not training and torch.is_grad_enabled() and x.requires_grad):
            y = super().forward(x)
            if residual is not None:
                y = y + residual
            return torch.relu(y) if relu else y
        momentum = 0.1 if self.momentum is None else self.momentum
        return _BNFn.apply(x, residual, self.weight, self.bias, self.running_mean, self.running_var,
                           self.num_batches_tracked if (training and self.track_running_stats) else None,
                           momentum, self.eps, training, relu)


# ---- fused forwards for the torchvision graphs the reference instantiates ------------------------------
def _basic_block_forward(self, x):
    identity = x
    out = self.bn1(self.conv1(x), relu=True)
    out = self.conv2(out)
    if self.downsample is not None:
        identity = self.downsample(x)
    return self.bn2(out, residual=identity, relu=True)


def _bottleneck_forward(self, x):
    identity = x
    out = self.bn1(self.conv1(x), relu=True)
    out = self.bn2(self.conv2(out), relu=True)
    out = self.conv3(out)
    if self.downsample is not None:
        identity = self.downsample(x)
    return self.bn3(out, residual=identity, relu=True)


def _resnet_forward_impl(self, x):
    x = self.bn1(self.conv1(x), relu=True)
    x = self.maxpool(x)
    x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
    x = torch.flatten(self.avgpool(x), 1)
    return self.fc(x)


class _FusedSeq(nn.Sequential):
    """nn.Sequential that runs `BatchNorm2dB200 -> ReLU` pairs as one fused call (VGG-BN features)."""

    def forward(self, x):
        mods = list(self)
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, BatchNorm2dB200) and i + 1 < len(mods) and isinstance(mods[i + 1], nn.ReLU):
                x = m(x, relu=True); i += 2
            else:
                x = m(x); i += 1
        return x


def convert_batchnorm(module: nn.Module):
    """Replace every nn.BatchNorm2d by BatchNorm2dB200 (same init: ones / zeros, no RNG consumed)."""
    for name, child in module.named_children():
        if type(child) is nn.BatchNorm2d:
            new = BatchNorm2dB200(child.num_features, eps=child.eps, momentum=child.momentum, affine=child.affine,
                                  track_running_stats=child.track_running_stats)
            new.load_state_dict(child.state_dict())
            new.train(child.training)
            setattr(module, name, new)
        else:
            convert_batchnorm(child)
    return module


def fuse_torchvision_blocks(net: nn.Module):
    from torchvision.models.resnet import BasicBlock, Bottleneck, ResNet
    from torchvision.models.vgg import VGG
    convert_batchnorm(net)
    for m in net.modules():
        if type(m) is BasicBlock:
            m.forward = types.MethodType(_basic_block_forward, m)
        elif type(m) is Bottleneck:
            m.forward = types.MethodType(_bottleneck_forward, m)
    if isinstance(net, ResNet) and isinstance(net.bn1, BatchNorm2dB200):
        net._forward_impl = types.MethodType(_resnet_forward_impl, net)
    if isinstance(net, VGG) and type(net.features) is nn.Sequential:
        net.features.__class__ = _FusedSeq
    return net
