"""Oracle: masked operators (test infrastructure only; see oracle/__init__.py).

Restates utils/mask_layers.py of the reference:
  * ConvMask.forward   -> ``masked_conv2d``   (utils/mask_layers.py:23-34)
  * LinearMask.forward -> ``masked_linear``   (utils/mask_layers.py:59-70)
  * Conv1dMask.forward -> ``masked_conv1d_k1``(utils/mask_layers.py:104-119)
and the autograd backward of those expressions (no reference source; derived from
``y = op(x, mask * w)``):  dX = op_dgrad(dY, mask*w), dW = mask * op_wgrad(x, dY),
db = sum(dY).

Everything here is torch-CPU fp32 (optionally with operands rounded to bf16 first,
which is what CUDA autocast feeds the tensor cores — base_harness.py:121-125).
"""
import torch
import torch.nn.functional as F


def _round_bf16(t: torch.Tensor) -> torch.Tensor:
    return t.to(torch.bfloat16).to(torch.float32)


def _operands(x, w, mask, bf16_operands):
    sparse_w = mask.to(w.dtype) * w  # mask_layers.py:25 — elementwise, fp32
    if bf16_operands:
        # autocast casts the (non-leaf) product and the activation to bf16; the
        # contraction accumulates in fp32.
        return _round_bf16(x.float()), _round_bf16(sparse_w.float())
    return x.float(), sparse_w.float()


def masked_conv2d(x, w, mask, bias=None, stride=1, padding=0, dilation=1, groups=1,
                  bf16_operands=False):
    """y = conv2d(x, mask*w, b, ...) — utils/mask_layers.py:23-34."""
    xe, we = _operands(x, w, mask, bf16_operands)
    return F.conv2d(xe, we, None if bias is None else bias.float(), stride, padding, dilation, groups)


def masked_linear(x, w, mask, bias=None, bf16_operands=False):
    """y = x @ (mask*w)^T + b — utils/mask_layers.py:69-70."""
    xe, we = _operands(x, w, mask, bf16_operands)
    return F.linear(xe, we, None if bias is None else bias.float())


def masked_conv1d_k1(x, w, mask, bias=None, bf16_operands=False):
    """Conv1dMask: w is [out, in, 1]; x [B, in] -> unsqueeze, conv1d(k=1), squeeze.

    utils/mask_layers.py:104-119.  Numerically a linear layer with w[:, :, 0].
    """
    xe, we = _operands(x, w, mask, bf16_operands)
    y = F.conv1d(xe.unsqueeze(-1), we, None if bias is None else bias.float())
    return y.squeeze(-1)


def masked_conv2d_grads(x, w, mask, dy, stride=1, padding=0, bf16_operands=False, has_bias=False):
    """(dX, dW, db) of ``masked_conv2d`` w.r.t. (x, w, bias) for upstream grad dy."""
    xe, we = _operands(x, w, mask, bf16_operands)
    dye = _round_bf16(dy.float()) if bf16_operands else dy.float()
    xe = xe.detach().requires_grad_(True)
    we = we.detach().requires_grad_(True)
    y = F.conv2d(xe, we, None, stride, padding)
    gx, gw = torch.autograd.grad(y, (xe, we), dye)
    dw = mask.float() * gw  # MulBackward of mask*w: grad flows to w scaled by mask
    db = dye.sum(dim=(0, 2, 3)) if has_bias else None
    return gx, dw, db


def masked_linear_grads(x, w2d, mask2d, dy, bf16_operands=False, has_bias=False):
    """(dX, dW, db) of ``masked_linear`` (also Conv1dMask with w[:, :, 0])."""
    xe, we = _operands(x, w2d, mask2d, bf16_operands)
    dye = _round_bf16(dy.float()) if bf16_operands else dy.float()
    x2 = xe.reshape(-1, xe.shape[-1])
    dy2 = dye.reshape(-1, dye.shape[-1])
    gx = (dy2 @ we).reshape(xe.shape)
    dw = mask2d.float() * (dy2.t() @ x2)
    db = dy2.sum(0) if has_bias else None
    return gx, dw, db
