"""Host-side helpers kept from the reference's ``utils/harness_utils.py`` (only what the hot path touches)."""
import os
import random
import uuid
from datetime import datetime

import numpy as np
import torch
import yaml


def set_seed(cfg, is_deterministic: bool = False) -> None:
    """np / random / torch / cuda seeds from experiment_params.seed (reference harness_utils.py:97-114)."""
    seed = cfg.experiment_params.seed
    np.random.seed(seed); random.seed(seed); torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)


def generate_densities(cfg, current_sparsity: float):
    """Density per level (reference harness_utils.py:117-145).  Float64 on the host on purpose:
    ``k = int((1 - density) * N)`` must come out identical.  Level 0 of iterative methods is dense."""
    method = cfg.pruning_params.prune_method
    target = cfg.pruning_params.target_sparsity
    if method in ("mag", "random_erk", "random_balanced"):
        rate = cfg.pruning_params.prune_rate
        out, cur, goal = [], 1 - current_sparsity, 1 - target
        while cur > goal:
            out.append(cur)
            cur *= 1 - rate
        if cur <= goal:
            out.append(cur)
        return out
    if method in ("er_erk", "er_balanced", "synflow", "snip"):
        return [1 - target]
    if method == "just dont":
        return [1.0]
    raise ValueError(f"Unknown pruning method: {method}")


def gen_expt_dir(cfg):
    """<base_dir>/<prefix>__<uuid6>__<time>/{checkpoints,metrics/level_wise_metrics,artifacts} (reference :49-94)."""
    prefix = (f"{cfg.dataset_params.dataset_name}_model_{cfg.model_params.model_name}"
              f"_trainingtype_{cfg.pruning_params.training_type}_prunemethod_{cfg.pruning_params.prune_method}"
              f"_target_{cfg.pruning_params.target_sparsity:.2f}_seed_{cfg.experiment_params.seed}"
              f"_budget_{cfg.experiment_params.epochs_per_level}epochs_lr_{cfg.optimizer_params.lr:.3f}"
              f"_mom_{cfg.optimizer_params.momentum:.1f}_wd_{cfg.optimizer_params.weight_decay:.4f}"
              f"_sched_{cfg.optimizer_params.scheduler_type}")
    expt_dir = os.path.join(cfg.experiment_params.base_dir,
                            f"{prefix}__{uuid.uuid4().hex[:6]}__{datetime.now().strftime('%Y%m%d_%H%M%S')}")
    for sub in ("checkpoints", "metrics/level_wise_metrics", "artifacts"):
        os.makedirs(os.path.join(expt_dir, sub), exist_ok=True)
    return prefix, expt_dir


def save_config(expt_dir, cfg) -> None:
    def plain(x):
        return {k: plain(v) for k, v in x.items()} if isinstance(x, dict) else x
    with open(os.path.join(expt_dir, "expt_config.yaml"), "w") as f:
        yaml.dump(plain(cfg), f, default_flow_style=False)


def unwrap(model):
    """PruneModel.model of a possibly wrapped model (reference save_model, harness_utils.py:354-365)."""
    m = getattr(model, "_orig_mod", model)
    m = getattr(m, "module", m)
    return m.model


# ---- device-resident checkpoint cache ------------------------------------------------------
# The level loop re-reads checkpoints it wrote moments earlier: model_level_{L-1}.pt at the start of level L and
# model_init.pt / model_rewind.pt at every rewind (run_experiment.py:96-105 of the reference: a torch.load + H2D of the
# whole state dict each time, 20+ levels).  The on-disk files and their format are unchanged; save_model additionally
# keeps a clone of what it wrote on the model's device, and load_checkpoint serves it as long as the file is still the
# one that was written (same size and mtime).  A few entries suffice: init / rewind / previous level.
_CKPT_CACHE = {}
_CKPT_CACHE_MAX = 4


def _file_key(path):
    st = os.stat(path)
    return (st.st_size, st.st_mtime_ns)


def load_checkpoint(path):
    """``torch.load(path)`` semantics, served from the device-resident cache when this process wrote the file."""
    ap = os.path.abspath(path)
    hit = _CKPT_CACHE.get(ap)
    if hit is not None:
        try:
            if _file_key(ap) == hit[0]:
                return hit[1]
        except OSError:
            pass
        _CKPT_CACHE.pop(ap, None)
    return torch.load(path)


def save_model(model, save_path, distributed: bool = False) -> None:
    """State dict of the INNER torchvision net: keys ``conv1.weight``, ``conv1.mask`` ... (checkpoint compatible)."""
    sd = unwrap(model).state_dict()
    torch.save(sd, save_path)
    try:
        ap = os.path.abspath(save_path)
        _CKPT_CACHE.pop(ap, None)
        while len(_CKPT_CACHE) >= _CKPT_CACHE_MAX:
            # evict the oldest per-level checkpoint first: model_init.pt / model_rewind.pt are re-read at every level
            levels = [k for k in _CKPT_CACHE if os.path.basename(k).startswith("model_level_")]
            _CKPT_CACHE.pop(levels[0] if levels else next(iter(_CKPT_CACHE)))
        _CKPT_CACHE[ap] = (_file_key(ap), {k: v.detach().clone() for k, v in sd.items()})
    except OSError:
        pass
    print(f"Model saved to {save_path}")
