"""Vision-transformer definitions for the DeiT configs (reference utils/deit.py:69-252).

The reference builds these from ``timm.models.vision_transformer.VisionTransformer``; timm is
not available offline, so this is a self-contained definition with the same block semantics
(pre-LN, fused qkv Linear with bias, scaled-dot-product attention, GELU MLP x4, class token,
learned position embedding, LayerNorm eps 1e-6, linear head).  Only ``nn.Linear`` layers are
masked (custom_models.CustomModel), the patch embedding conv and the attention matmuls are not.
"""
from functools import partial

import torch
import torch.nn as nn
import torch.nn.functional as F


class _Attention(nn.Module):
    def __init__(self, dim, heads, qkv_bias=True):
        super().__init__()
        self.heads = heads
        self.qkv = nn.Linear(dim, dim * 3, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        b, t, c = x.shape
        qkv = self.qkv(x).reshape(b, t, 3, self.heads, c // self.heads).permute(2, 0, 3, 1, 4)
        out = F.scaled_dot_product_attention(qkv[0], qkv[1], qkv[2])
        return self.proj(out.transpose(1, 2).reshape(b, t, c))


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class _Block(nn.Module):
    def __init__(self, dim, heads, mlp_ratio, qkv_bias, norm_layer):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = _Attention(dim, heads, qkv_bias)
        self.norm2 = norm_layer(dim)
        self.mlp = _Mlp(dim, int(dim * mlp_ratio))

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


class _PatchEmbed(nn.Module):
    def __init__(self, img_size, patch, in_chans, dim):
        super().__init__()
        self.num_patches = (img_size // patch) ** 2
        self.proj = nn.Conv2d(in_chans, dim, kernel_size=patch, stride=patch)

    def forward(self, x):
        return self.proj(x).flatten(2).transpose(1, 2)


class VisionTransformer(nn.Module):
    def __init__(self, img_size=224, patch_size=16, in_chans=3, num_classes=1000, embed_dim=768, depth=12,
                 num_heads=12, mlp_ratio=4.0, qkv_bias=True, norm_layer=None, distilled=False):
        super().__init__()
        norm_layer = norm_layer or partial(nn.LayerNorm, eps=1e-6)
        self.patch_embed = _PatchEmbed(img_size, patch_size, in_chans, embed_dim)
        extra = 2 if distilled else 1
        self.cls_token = nn.Parameter(torch.zeros(1, 1, embed_dim))
        self.dist_token = nn.Parameter(torch.zeros(1, 1, embed_dim)) if distilled else None
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches + extra, embed_dim))
        self.blocks = nn.ModuleList([_Block(embed_dim, num_heads, mlp_ratio, qkv_bias, norm_layer) for _ in range(depth)])
        self.norm = norm_layer(embed_dim)
        self.head = nn.Linear(embed_dim, num_classes)
        self.head_dist = nn.Linear(embed_dim, num_classes) if distilled else None
        nn.init.trunc_normal_(self.pos_embed, std=0.02)
        nn.init.trunc_normal_(self.cls_token, std=0.02)
        if distilled:
            nn.init.trunc_normal_(self.dist_token, std=0.02)

    def forward(self, x):
        x = self.patch_embed(x)
        toks = [self.cls_token.expand(x.shape[0], -1, -1)]
        if self.dist_token is not None:
            toks.append(self.dist_token.expand(x.shape[0], -1, -1))
        x = torch.cat(toks + [x], dim=1) + self.pos_embed
        for blk in self.blocks:
            x = blk(x)
        x = self.norm(x)
        if self.head_dist is None:
            return self.head(x[:, 0])
        a, b = self.head(x[:, 0]), self.head_dist(x[:, 1])
        return (a, b) if self.training else (a + b) / 2


def _deit(dim, heads, img=224, distilled=False):
    return VisionTransformer(img_size=img, patch_size=16, embed_dim=dim, depth=12, num_heads=heads, mlp_ratio=4,
                             qkv_bias=True, norm_layer=partial(nn.LayerNorm, eps=1e-6), distilled=distilled)


def local_deit_tiny_patch16_224(**kw):
    return _deit(192, 3)


def local_deit_small_patch16_224(**kw):
    return _deit(384, 6)


def local_deit_base_patch16_224(**kw):
    return _deit(768, 12)


def local_deit_small_distilled_patch16_224(**kw):
    return _deit(384, 6, distilled=True)


def local_deit_base_distilled_patch16_224(**kw):
    return _deit(768, 12, distilled=True)


def local_deit_base_patch16_384(**kw):
    return _deit(768, 12, img=384)


def local_deit_base_distilled_patch16_384(**kw):
    return _deit(768, 12, img=384, distilled=True)
