"""Oracle: pruning score / threshold / mask (test infrastructure only).

numpy restatement of utils/pruning_utils.py of the reference, bit-exact by
construction: every score is a short chain of IEEE fp32 multiplies followed by
``abs``; the threshold is an exact order statistic; the mask is an exact compare.

  score_mag        utils/pruning_utils.py:75        |mask * w|
  score_grad       utils/pruning_utils.py:190 (snip: (g*w)*mask), :267 (synflow: (mask*g)*w)
  kth_smallest     torch.kthvalue semantics used at :79,:137,:195,:275,:337
                   (k is 1-indexed, NaN sorts above +inf, k==0 raises)
  apply_threshold  utils/pruning_utils.py:84-87     where(score <= thr, 0, 1)
  prune_global     prune_mag :61-89, prune_snip :186-203, prune_synflow :263-283
  prune_per_layer  prune_random_erk :129-144, prune_random_balanced :329-345
  erk_keep_probabilities      :357-371 (fp32 tensor arithmetic, restated with torch)
  balanced_keep_probabilities :388-407 (Python float arithmetic)
  count_zeros / overall_sparsity_percent   utils/custom_models.py:51-62
  generate_densities           utils/harness_utils.py:117-145
"""
import numpy as np

SCORE_MAG = 0      # |m*w|
SCORE_SNIP = 1     # |(g*w)*m|
SCORE_SYNFLOW = 2  # |(m*g)*w|


def _f32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def sortable_key(x):
    """uint32 key with the total order torch's radix select uses.

    ATen SortingRadixSelect.cuh:20-39 (TopKTypeConfig<float>::convert): flip all bits
    of negatives, set the sign bit of non-negatives, and map every NaN to 0xFFFFFFFF
    (largest).
    """
    x = _f32(x)
    u = x.view(np.uint32)
    sign = u >> np.uint32(31)
    flip = (np.uint32(0) - sign) | np.uint32(0x80000000)   # 0xFFFFFFFF for negatives, 0x80000000 otherwise
    key = u ^ flip
    nan = (u & np.uint32(0x7FFFFFFF)) > np.uint32(0x7F800000)
    if nan.any():
        key = key.copy()
        key[nan] = np.uint32(0xFFFFFFFF)
    return key


def kth_smallest(scores, k):
    """Value of the k-th smallest element (1-indexed) — torch.kthvalue semantics."""
    scores = _f32(scores).reshape(-1)
    n = scores.size
    if k < 1 or k > n:
        # torch: "kthvalue(): selected number k out of range for dimension 0"
        raise RuntimeError(f"kthvalue(): selected number k out of range for dimension 0 (k={k}, n={n})")
    keys = sortable_key(scores)
    kk = np.partition(keys, k - 1)[k - 1]
    if kk == np.uint32(0xFFFFFFFF):
        return np.float32(np.nan)
    u = np.uint32(kk)
    u = (u ^ np.uint32(0x80000000)) if (u & np.uint32(0x80000000)) else np.uint32(~u)
    return np.array([u], dtype=np.uint32).view(np.float32)[0]


def score_mag(w, m):
    """|mask * w| in fp32 — pruning_utils.py:75 (also :112 with w := randn)."""
    return np.abs(_f32(m) * _f32(w))


def score_grad(w, g, m, kind):
    """SNIP / SynFlow saliency in the reference's evaluation order."""
    w, g, m = _f32(w), _f32(g), _f32(m)
    with np.errstate(invalid="ignore", over="ignore"):
        if kind == SCORE_SNIP:
            return np.abs((g * w) * m)      # pruning_utils.py:190
        if kind == SCORE_SYNFLOW:
            return np.abs((m * g) * w)      # pruning_utils.py:267
    raise ValueError(kind)


def layer_scores(ws, ms, gs=None, kind=SCORE_MAG):
    if kind == SCORE_MAG:
        return [score_mag(w, m) for w, m in zip(ws, ms)]
    return [score_grad(w, g, m, kind) for w, g, m in zip(ws, gs, ms)]


def global_threshold(scores, density):
    """k = int((1 - density) * N) in Python float64; thr = kthvalue(cat(scores), k)."""
    n = int(sum(s.size for s in scores))
    k = int((1 - density) * n)              # pruning_utils.py:78 — host float64
    flat = np.concatenate([s.reshape(-1) for s in scores])  # :77
    return kth_smallest(flat, k), k          # raises for k == 0 like the reference (:79 before :81)


def apply_threshold(score, thr):
    """mask = where(score <= thr, 0., 1.) — pruning_utils.py:84-87 (ties pruned)."""
    with np.errstate(invalid="ignore"):
        return np.where(_f32(score) <= np.float32(thr), np.float32(0.0), np.float32(1.0)).astype(np.float32)


def prune_global(ws, ms, density, gs=None, kind=SCORE_MAG):
    """New masks for prune_mag / prune_snip / prune_synflow. Returns (masks, thr, k)."""
    scores = layer_scores(ws, ms, gs, kind)
    thr, k = global_threshold(scores, density)
    return [apply_threshold(s, thr) for s in scores], thr, k


def prune_per_layer(noises, ms, keep_fracs):
    """prune_random_erk / prune_random_balanced after the RNG draw.

    ``noises`` are the per-layer ``randn_like`` draws (kept in torch for Philox
    parity, pruning_utils.py:112,314); ``keep_fracs`` the per-layer keep fractions
    (fp32 tensors for erk, Python floats for balanced).  k_l = int((1-p_l) * n_l);
    k_l == 0 -> threshold 0 (:134-135,:334-335).
    """
    out, ks = [], []
    for z, m, p in zip(noises, ms, keep_fracs):
        s = score_mag(z, m)
        # (1 - p) * numel evaluated the way python/torch does: p may be a 0-dim fp32
        # tensor (erk) -> fp32 arithmetic then int(); or a float (balanced) -> float64.
        k = int((1 - p) * s.size)
        thr = np.float32(0.0) if k == 0 else kth_smallest(s, k)
        out.append(apply_threshold(s, thr))
        ks.append(k)
    return out, ks


def count_zeros(ms):
    return int(sum(int((np.asarray(m) == 0).sum()) for m in ms))


def overall_sparsity_percent(ms):
    """custom_models.py:51-62 — returns PERCENT."""
    total = int(sum(np.asarray(m).size for m in ms))
    return (count_zeros(ms) / total) * 100 if total > 0 else 0


def erk_keep_probabilities(shapes, density):
    """ERK keep-probabilities, pruning_utils.py:357-371 (identical in :117-127).

    The reference does this arithmetic with torch fp32 tensors; restated with the
    same torch ops so the values are bit-identical 0-dim fp32 tensors.
    """
    import torch
    sparsity_list, num_params_list, total = [], [], 0
    for shp in shapes:
        numel = int(np.prod(shp))
        sparsity_list.append(torch.tensor(tuple(shp)).sum() / numel)   # :362
        num_params_list.append(numel)
        total += numel
    kept = (torch.tensor(sparsity_list) * torch.tensor(num_params_list)).sum()  # :366-368
    c = (total * density) / kept                                              # :369-370
    return [torch.clamp(c * s, 0, 1) for s in sparsity_list]                   # :371


def balanced_keep_probabilities(numels, density):
    """Balanced keep-probabilities, pruning_utils.py:388-407 (and :298-327). Python floats."""
    total = sum(numels)
    L = len(numels)
    X = density * total / L
    out = []
    for l, n in enumerate(numels):
        if X / n < 1.0:
            out.append(X / n)
        else:
            out.append(1)
            diff = X - n
            X = X + diff / (L - l)
    return out


def generate_densities(prune_method, target_sparsity, prune_rate=0.2, current_sparsity=0.0):
    """utils/harness_utils.py:117-145 (float64 host arithmetic; level 0 is dense for iterative)."""
    if prune_method in ("mag", "random_erk", "random_balanced"):
        out = []
        cur = 1 - current_sparsity
        tgt = 1 - target_sparsity
        while cur > tgt:
            out.append(cur)
            cur *= 1 - prune_rate
        if cur <= tgt:
            out.append(cur)
        return out
    if prune_method in ("er_erk", "er_balanced", "synflow", "snip"):
        return [1 - target_sparsity]
    if prune_method == "just dont":
        return [1.0]
    raise ValueError(f"Unknown pruning method: {prune_method}")
