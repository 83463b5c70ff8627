"""Configuration with the reference's surface: `general.*`, `data.*`, `model.*`, `opt.*`,
`logging.*`, attribute access (`cfg.data.fov`, `hasattr(cfg.data, "training_resolution")`).

hydra / omegaconf are not available here, so this module provides
  * `Cfg`: a small attribute-access mapping (OmegaConf-like `cfg.a.b`, `in`, `hasattr`, `to_dict`);
  * `compose(config_name, config_dir=None, overrides=())`: hydra-style composition of a YAML tree.  With
    `config_dir=None` the defaults below are used -- they restate the values of the reference's
    `configs/settings.yaml`, `configs/dataset/shapenet.yaml` and `configs/transformer_pretraining.yaml`; pointing
    `config_dir` at the reference's own `configs/` directory composes those files directly
    (`defaults:` lists with `/group@_here_` entries and `_self_`, missing optional groups ignored);
  * `key=value` dotted overrides as on the hydra command line.
"""
from __future__ import annotations

import copy
import os
from typing import Any, Dict, Iterable, Optional

import yaml


class Cfg(dict):
    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError:
            raise AttributeError(k)
        return v

    def __setattr__(self, k, v):
        self[k] = _wrap(v)

    def __deepcopy__(self, memo):
        return Cfg({k: copy.deepcopy(v, memo) for k, v in self.items()})

    def to_dict(self) -> Dict[str, Any]:
        return {k: (v.to_dict() if isinstance(v, Cfg) else v) for k, v in self.items()}


def _wrap(v):
    if isinstance(v, dict) and not isinstance(v, Cfg):
        return Cfg({k: _wrap(x) for k, x in v.items()})
    if isinstance(v, list):
        return [_wrap(x) for x in v]
    return v


def _merge(dst: dict, src: dict) -> dict:
    for k, v in src.items():
        if isinstance(v, dict) and isinstance(dst.get(k), dict):
            _merge(dst[k], v)
        else:
            dst[k] = copy.deepcopy(v)
    return dst


_SETTINGS = {
    "opt": {"betas": [0.9, 0.999], "imgs_per_obj": 4,
            "ema": {"use": True, "update_every": 10, "update_after_step": 100, "beta": 0.9999},
            "lambda_lpips": 0.01, "pretrained_ckpt": None, "record_img": True},
    "model": {"max_sh_degree": 1, "isotropic": False, "image_branch": "stem", "vae_weights": None},
    "logging": {"ckpt_iterations": 2000, "val_log": 2000, "loss_log": 10, "loop_log": 2000, "render_log": 2000,
                "centered": True},
}
_SHAPENET = {"data": {"znear": 0.5, "zfar": 2, "category": "shapenet", "white_background": False,
                      "dataset_root": "/path/to/shapenet/dataset"}}
_TRANSFORMER = {
    "general": {"cuda": True, "device": 0, "random_seed": 42},
    "data": {"fov": 49.13434264120263, "training_resolution": 128, "input_images": 1},
    "model": {"backbone_type": "transformer", "in_channels": 3, "aug": True, "offset_scale": 1.0},
    "opt": {"iterations": 100000, "mode": "train", "level": "object", "use_fusion": True, "base_lr": 0.0001,
            "batch_size": 32, "test_generation_num": 1, "loss": "focal_l2", "non_bg_color_loss_rate": 4,
            "bg_color_loss_rate": 1, "step_lr": 20000, "lr_gamma": 0.8, "start_lpips_after": 50000},
}
_POINTMLP = copy.deepcopy(_TRANSFORMER)
_POINTMLP["model"].update({"backbone_type": "pointmlp", "in_channels": 4})
# configs/dataset/scannet.yaml + configs/sparseunet_pretraining.yaml / ptv3_pretraining.yaml of the reference
_SCANNET = {"data": {"znear": 0.2, "zfar": 10, "category": "scannet", "white_background": True, "use_neighbor_imgs": True,
                     "use_self_supervise": False, "supervised_max_distance": 5,
                     "pts_dataset_root": "/path/to/scannet/dataset", "rgb_dataset_root": "/path/to/scannet/color"}}
_SPARSEUNET = {
    "general": {"cuda": True, "device": [0, 1, 2, 3], "random_seed": 42},
    "data": {"fov": 57.9516132895, "training_width": 160, "training_height": 120, "input_images": 8},
    "model": {"backbone_type": "sparseunet", "in_channels": 3, "aug": False, "offset_scale": 0.2},
    "opt": {"iterations": 60000, "mode": "train", "level": "scene", "use_fusion": True, "base_lr": 0.0001, "batch_size": 4,
            "test_generation_num": 1, "loss": "l2", "step_lr": 10000, "lr_gamma": 0.9, "start_lpips_after": 30000},
}
_PTV3 = copy.deepcopy(_SPARSEUNET)
_PTV3["model"].update({"backbone_type": "ptv3"})
BUILTIN = {
    "settings": _SETTINGS,
    "dataset/shapenet": _SHAPENET,
    "transformer_pretraining": {"defaults": ["/settings@_here_", "/dataset/shapenet@_here_"], **_TRANSFORMER},
    "pointmlp_pretraining": {"defaults": ["/settings@_here_", "/dataset/shapenet@_here_"], **_POINTMLP},
    "dataset/scannet": _SCANNET,
    "sparseunet_pretraining": {"defaults": ["/settings@_here_", "/dataset/scannet@_here_"], **_SPARSEUNET},
    "ptv3_pretraining": {"defaults": ["/settings@_here_", "/dataset/scannet@_here_"], **_PTV3},
    "default_config": {"defaults": ["/transformer_pretraining@_here_"]},
}


def _load_node(name: str, config_dir: Optional[str]) -> Optional[dict]:
    name = name.strip("/")
    if config_dir is None:
        node = BUILTIN.get(name)
        return copy.deepcopy(node) if node is not None else None
    path = os.path.join(config_dir, name + ".yaml")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        docs = [d for d in yaml.safe_load_all(f) if d is not None]
    return docs[0] if docs else {}


def _compose_node(name: str, config_dir: Optional[str], depth: int = 0) -> dict:
    node = _load_node(name, config_dir)
    if node is None:
        return {}
    if depth > 16:
        raise RuntimeError("config defaults nest too deeply")
    defaults = node.pop("defaults", []) or []
    out: dict = {}
    self_done = False
    for d in defaults:
        if d == "_self_":
            _merge(out, node)
            self_done = True
            continue
        if isinstance(d, dict):          # "group: option" form (wandb/hydra/cam_embd in settings.yaml) -> optional
            for grp, opt in d.items():
                sub = _compose_node(f"{grp}/{opt}", config_dir, depth + 1)
                if sub:
                    _merge(out, {grp.strip("/"): sub})
            continue
        target = str(d).split("@")[0]
        _merge(out, _compose_node(target, config_dir, depth + 1))
    if not self_done:
        _merge(out, node)
    return out


def _parse_value(s: str):
    try:
        return yaml.safe_load(s)
    except Exception:
        return s


def compose(config_name: str = "default_config", config_dir: Optional[str] = None,
            overrides: Iterable[str] = ()) -> Cfg:
    cfg = _compose_node(config_name, config_dir)
    if not cfg:
        raise FileNotFoundError(f"config {config_name!r} not found" + (f" under {config_dir}" if config_dir else ""))
    for ov in overrides:
        key, _, val = ov.partition("=")
        cur = cfg
        parts = key.lstrip("+").split(".")
        for p in parts[:-1]:
            cur = cur.setdefault(p, {})
        cur[parts[-1]] = _parse_value(val)
    cfg.setdefault("general", {})
    dev = cfg["general"].get("device", 0)
    # train_network.py:562-567: multiple_gpu is derived from the device list at run time
    cfg["general"]["multiple_gpu"] = bool(hasattr(dev, "__len__") and not isinstance(dev, str) and len(dev) > 1)
    return _wrap(cfg)
