"""Oracle: the data path either side of the model (test infrastructure only).

  batch_crop / batch_flip_lr / batch_cutout / augment
        utils/dataset.py:38-98 of the reference (airbench-style GPU augmentation used by CifarLoader.__iter__,
        :192-226), restated as pure index arithmetic on numpy arrays for GIVEN random draws — the draws themselves
        (torch.randint / torch.rand on the images' device, in the order crop shifts -> flip mask -> cutout corners)
        stay torch's in the product, so parity is "same draws in -> same pixels out", bit-exact.
  philox4x32_10 / synth_normal / synth_labels
        the synthetic generator that stands in for the data sets (none are reachable here): Philox4x32-10
        (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11 — multipliers 0xD2511F53 /
        0xCD9E8D57, Weyl key increments 0x9E3779B9 / 0xBB67AE85, ten rounds) with counter (i, 0, 0, 0) and the 64-bit
        seed as key; normals by Box-Muller on 24-bit uniforms.  No reference counterpart (FFCV / CIFAR tensors are data).
"""
import numpy as np


def batch_crop(images, crop_size, shifts):
    """utils/dataset.py:43-69: images [N,C,Hp,Wp] (reflect-padded), shifts [N,2] in [-r, r]: out = window at (r+sy, r+sx)."""
    images = np.asarray(images)
    n = images.shape[0]
    r = (images.shape[-1] - crop_size) // 2
    out = np.empty((n, images.shape[1], crop_size, crop_size), images.dtype)
    for i in range(n):
        sy, sx = int(shifts[i][0]), int(shifts[i][1])
        out[i] = images[i, :, r + sy:r + sy + crop_size, r + sx:r + sx + crop_size]
    return out


def batch_flip_lr(inputs, flip_mask):
    """utils/dataset.py:38-40: where(flip_mask, inputs.flip(-1), inputs)."""
    inputs = np.asarray(inputs)
    m = np.asarray(flip_mask).astype(bool).reshape(-1, 1, 1, 1)
    return np.where(m, inputs[..., ::-1], inputs)


def batch_cutout(inputs, size, corner_y, corner_x):
    """utils/dataset.py:72-98: zero the size x size square whose top-left corner is (corner_y, corner_x)."""
    inputs = np.asarray(inputs)
    n, c, h, w = inputs.shape
    yy = np.arange(h).reshape(1, 1, h, 1) - np.asarray(corner_y).reshape(-1, 1, 1, 1)
    xx = np.arange(w).reshape(1, 1, 1, w) - np.asarray(corner_x).reshape(-1, 1, 1, 1)
    mask = ((yy >= 0) & (yy < size)) & ((xx >= 0) & (xx < size))
    return np.where(mask, np.zeros((), inputs.dtype), inputs)


def augment(padded, crop_size, shifts=None, flip_mask=None, cutout=0, corner_y=None, corner_x=None):
    """CifarLoader.__iter__ (:204-221) for one epoch: translate -> flip -> cutout, each optional."""
    x = batch_crop(padded, crop_size, shifts) if shifts is not None else np.asarray(padded)
    if flip_mask is not None:
        x = batch_flip_lr(x, flip_mask)
    if cutout:
        x = batch_cutout(x, cutout, corner_y, corner_x)
    return x


_M0, _M1, _W0, _W1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)


def philox4x32_10(counters, seed):
    """uint32 [len(counters), 4]: Philox4x32-10 of counter (ctr_lo, ctr_hi, 0, 0) under key (seed_lo, seed_hi)."""
    ctr = np.asarray(counters, dtype=np.uint64)
    c0 = (ctr & np.uint64(0xFFFFFFFF)).astype(np.uint32); c1 = (ctr >> np.uint64(32)).astype(np.uint32)
    c2 = np.zeros_like(c0); c3 = np.zeros_like(c0)
    k0 = np.uint32(seed & 0xFFFFFFFF); k1 = np.uint32((seed >> 32) & 0xFFFFFFFF)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = _M0 * c0.astype(np.uint64); p1 = _M1 * c2.astype(np.uint64)
            hi0 = (p0 >> np.uint64(32)).astype(np.uint32); lo0 = (p0 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
            hi1 = (p1 >> np.uint64(32)).astype(np.uint32); lo1 = (p1 & np.uint64(0xFFFFFFFF)).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0 = np.uint32((int(k0) + int(_W0)) & 0xFFFFFFFF); k1 = np.uint32((int(k1) + int(_W1)) & 0xFFFFFFFF)
    return np.stack([c0, c1, c2, c3], axis=1)


def synth_words(n, seed, offset=0):
    q = (n + 3) // 4
    return philox4x32_10(np.arange(q, dtype=np.uint64) + np.uint64(offset), seed).reshape(-1)[:n]


def synth_normal(n, seed, offset=0):
    """N(0,1) fp32: Box-Muller on pairs of words, u1 = ((w >> 8) + 1) / 2^24 in (0, 1], u2 = (w >> 8) / 2^24."""
    w = synth_words((n + 3) // 4 * 4, seed, offset).reshape(-1, 2)
    u1 = ((w[:, 0] >> np.uint32(8)).astype(np.float32) + np.float32(1.0)) * np.float32(1.0 / 16777216.0)
    u2 = (w[:, 1] >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    rad = np.sqrt(np.float32(-2.0) * np.log(u1.astype(np.float64))).astype(np.float32)
    ang = 2.0 * np.pi * u2.astype(np.float64)
    out = np.stack([rad * np.cos(ang).astype(np.float32), rad * np.sin(ang).astype(np.float32)], axis=1).reshape(-1)
    return out[:n].astype(np.float32)


def synth_labels(n, num_classes, seed, offset=0):
    w = synth_words(n, seed, offset).astype(np.uint64)
    return ((w * np.uint64(num_classes)) >> np.uint64(32)).astype(np.int64)
