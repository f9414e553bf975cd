"""Hydra-free composer for the reference's ``conf/`` tree (hydra / omegaconf are not installed here).

Understands exactly what the reference uses (run_experiment.py:21, README.md:85-91): a top-level YAML with a
``defaults:`` list selecting one file per config group, ``--config-name=<name>``, ``group=file`` to swap a group,
``a.b=c`` / ``+a.b=c`` value overrides.  The result is a dict with attribute access (enough of a DictConfig for the
harness and pruning code).  Point ``conf_dir`` at the reference's own ``conf/`` directory to consume it unchanged.
"""
import os
import sys

import yaml


class Cfg(dict):
    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return v

    def __setattr__(self, k, v):
        self[k] = v


def _wrap(x):
    if isinstance(x, dict):
        return Cfg({k: _wrap(v) for k, v in x.items()})
    if isinstance(x, list):
        return [_wrap(v) for v in x]
    return x




def _fix_floats(node):
    """PyYAML reads '5e-4' / '1e-4' (no dot) as strings; hydra/omegaconf read them as floats."""
    if isinstance(node, dict):
        return {k: _fix_floats(v) for k, v in node.items()}
    if isinstance(node, list):
        return [_fix_floats(v) for v in node]
    if isinstance(node, str):
        try:
            return float(node) if any(c in node for c in "eE.") and node.replace("-", "").replace("+", "").replace(".", "").replace("e", "").replace("E", "").isdigit() else node
        except ValueError:
            return node
    return node


def _parse_scalar(text):
    """CLI override value: YAML scalar rules plus hydra's float forms ('5e-4', '1e-1' are floats, not strings)."""
    return _fix_floats(yaml.safe_load(text)) if text != "" else ""


def compose(config_name, overrides=(), conf_dir="conf"):
    top = yaml.safe_load(open(os.path.join(conf_dir, f"{config_name}.yaml"))) or {}
    groups, order = {}, []
    for item in top.pop("defaults", []):
        if item == "_self_":
            continue
        if isinstance(item, dict):
            (g, name), = item.items()
            groups[g] = name; order.append(g)
    values = []
    for ov in overrides:
        key, _, val = ov.partition("=")
        plus = key.startswith("+")
        key = key.lstrip("+")
        if "." not in key and (key in groups or os.path.isdir(os.path.join(conf_dir, key))):
            if key not in groups:
                order.append(key)
            groups[key] = val                                  # group selection, e.g. pruning_params=iterative_imp
        else:
            values.append((key, _parse_scalar(val), plus))
    cfg = dict(top)
    for g in order:
        path = os.path.join(conf_dir, g, f"{groups[g]}.yaml")
        cfg[g] = _fix_floats(yaml.safe_load(open(path)) or {})
    for key, val, plus in values:
        node = cfg
        parts = key.split(".")
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        if not plus and parts[-1] not in node:
            raise KeyError(f"override '{key}' does not exist in the config (use +{key}=... to add it)")
        node[parts[-1]] = val
    return _wrap(cfg)


def parse_cli(argv=None):
    """``--config-name=x [--config-path=dir] k=v ...`` -> (config_name, conf_dir, overrides)."""
    argv = list(sys.argv[1:] if argv is None else argv)
    name, conf_dir, ov = "config", None, []
    for a in argv:
        if a.startswith("--config-name"):
            name = a.split("=", 1)[1]
        elif a.startswith("--config-path"):
            conf_dir = a.split("=", 1)[1]
        else:
            ov.append(a)
    return name, conf_dir, ov
