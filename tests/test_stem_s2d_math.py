"""CPU: the index algebra of the round-2 stem kernel ("space-to-depth stem").

A 7x7 / stride-2 / pad-3 convolution over a 3-channel image equals a 4x4 / stride-1 convolution over the 2x2
space-to-depth image (12 channels) with the filter taps re-indexed; the implicit-GEMM kernel could then read the
input with 16-channel (32-byte) im2col TMA rows and no materialised matrix.  This test pins the mapping
(tap (r, s) -> block offset (r + 1) // 2, parity (r + 1) % 2) so the kernel work can start from a checked formula.
"""
import torch
import torch.nn.functional as F


def space_to_depth(x, pad):
    """[N, C, H, W] -> [N, 4C, (H + 2 pad) / 2, (W + 2 pad) / 2]; channel index (dh * 2 + dw) * C + c."""
    x = F.pad(x, (pad, pad, pad, pad))
    n, c, h, w = x.shape
    x = x.view(n, c, h // 2, 2, w // 2, 2).permute(0, 3, 5, 1, 2, 4)
    return x.reshape(n, 4 * c, h // 2, w // 2)


def s2d_filter(w):
    """[Cout, C, 7, 7] (stride 2, pad 3) -> [Cout, 4C, 4, 4] (stride 1 over the pad-4 space-to-depth image)."""
    cout, c, k, _ = w.shape
    w2 = torch.zeros(cout, 4 * c, 4, 4, dtype=w.dtype)
    for r in range(k):
        for s in range(k):
            br, dh = (r + 1) // 2, (r + 1) % 2          # image padded by 4: input row 2p - 4 + (r + 1)
            bs, dw = (s + 1) // 2, (s + 1) % 2
            w2[:, (dh * 2 + dw) * c:(dh * 2 + dw + 1) * c, br, bs] = w[:, :, r, s]
    return w2


def test_7x7_stride2_equals_4x4_over_space_to_depth():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 3, 32, 32, generator=g, dtype=torch.float64)
    w = torch.randn(8, 3, 7, 7, generator=g, dtype=torch.float64)
    ref = F.conv2d(x, w, stride=2, padding=3)
    xs = space_to_depth(x, 4)                            # pad 4 = pad 3 + one more row/column so that blocks align
    got = F.conv2d(xs, s2d_filter(w))
    assert got.shape[-1] == ref.shape[-1] + 1
    assert torch.allclose(got[:, :, :ref.shape[2], :ref.shape[3]], ref, atol=1e-10)
