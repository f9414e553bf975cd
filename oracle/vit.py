"""Oracle: DeiT vision transformer with masked Linears on CPU (test infrastructure only).

The reference builds DeiT from ``timm.models.vision_transformer.VisionTransformer`` with the
hyper-parameters in utils/deit.py:69-112 (patch 16, depth 12, mlp_ratio 4, qkv_bias=True,
LayerNorm eps 1e-6; tiny/small/base = 192/384/768 dims with 3/6/12 heads) and then replaces every
``nn.Linear`` by ``LinearMask`` (utils/custom_models.py:241-245, utils/mask_layers.py:55-70:
``F.linear(x, mask * w, b)``).  timm 0.x is not vendored under /root/reference and is absent from
this image (requirements.txt lists the bare name, no pin), so the block semantics are restated
here from timm's published ``VisionTransformer.forward_features`` / ``Block`` / ``Attention`` /
``Mlp``:

    x  = patch_embed(img)                       Conv2d(3, D, k=16, s=16) -> [B, 196, D]
    x  = cat(cls_token, x) + pos_embed          [B, 197, D]
    for each block:  x = x + proj(softmax(q k^T / sqrt(d)) v)   with q,k,v = split(qkv(LN(x)))
                     x = x + fc2(gelu(fc1(LN(x))))
    logits = head(LN(x)[:, 0])

Written as plain functions over a flat parameter dict (no module classes shared with the product);
parameter names follow timm's state-dict keys so weights move between product, oracle and a timm
checkpoint.  "parity unpinned" against timm itself: no golden vector of timm's output exists here;
the product's own ViT (turboprune_b200/utils/vit.py) is pinned against this restatement only.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

_SIZES = {"local_deit_tiny_patch16_224": (192, 3), "local_deit_small_patch16_224": (384, 6),
          "local_deit_base_patch16_224": (768, 12)}


class RefMaskedLinear(nn.Linear):
    """utils/mask_layers.py:55-70 — weight [out, in] times an fp32 ``mask`` buffer of the same shape."""

    def __init__(self, in_features, out_features, bias=True):
        super().__init__(in_features, out_features, bias=bias)
        self.register_buffer("mask", torch.ones_like(self.weight))

    def forward(self, x):
        return F.linear(x, self.mask.to(self.weight.device) * self.weight, self.bias)


class _Holder(nn.Module):
    """Namespace module: gives nested state-dict keys (``blocks.3.attn.qkv.weight``) without defining behaviour."""


class OracleDeiT(nn.Module):
    def __init__(self, name="local_deit_small_patch16_224", num_classes=1000, img=224, patch=16, depth=12):
        super().__init__()
        dim, heads = _SIZES[name]
        self.dim, self.heads, self.depth = dim, heads, depth
        pe = _Holder(); pe.proj = nn.Conv2d(3, dim, kernel_size=patch, stride=patch)
        self.patch_embed = pe
        n_tok = (img // patch) ** 2 + 1
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.zeros(1, n_tok, dim))
        blocks = []
        for _ in range(depth):
            b = _Holder()
            b.norm1 = nn.LayerNorm(dim, eps=1e-6)
            b.attn = _Holder(); b.attn.qkv = RefMaskedLinear(dim, 3 * dim); b.attn.proj = RefMaskedLinear(dim, dim)
            b.norm2 = nn.LayerNorm(dim, eps=1e-6)
            b.mlp = _Holder(); b.mlp.fc1 = RefMaskedLinear(dim, 4 * dim); b.mlp.fc2 = RefMaskedLinear(4 * dim, dim)
            blocks.append(b)
        self.blocks = nn.ModuleList(blocks)
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.head = RefMaskedLinear(dim, num_classes)

    def forward(self, img):
        B = img.shape[0]
        x = self.patch_embed.proj(img)                                  # [B, D, 14, 14]
        x = x.reshape(B, self.dim, -1).permute(0, 2, 1)                 # [B, 196, D]
        x = torch.cat([self.cls_token.expand(B, 1, self.dim), x], dim=1) + self.pos_embed
        hd = self.dim // self.heads
        for b in self.blocks:
            y = b.norm1(x)
            qkv = b.attn.qkv(y).reshape(B, -1, 3, self.heads, hd)
            q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3))   # [B, heads, T, hd]
            att = torch.softmax((q @ k.transpose(-1, -2)) * (1.0 / math.sqrt(hd)), dim=-1)
            y = (att @ v).permute(0, 2, 1, 3).reshape(B, -1, self.dim)
            x = x + b.attn.proj(y)
            y = b.mlp.fc1(b.norm2(x))
            x = x + b.mlp.fc2(F.gelu(y))
        return self.head(self.norm(x)[:, 0])


def build(name="local_deit_small_patch16_224"):
    return OracleDeiT(name)


def masked_layers(net):
    return [(n, m) for n, m in net.named_modules() if isinstance(m, RefMaskedLinear)]
