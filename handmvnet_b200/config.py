"""YAML -> cfg dict with the reference's schema and derived keys (reference src/config.py:35-51).

Unlike the reference module, nothing is parsed at import time: call load_config(path).
"""
import os

import yaml


def load_config(yaml_file: str) -> dict:
    if not os.path.exists(yaml_file):
        raise FileNotFoundError(f"[Error] config file doesn't exist: {yaml_file}")
    with open(yaml_file, "r") as f:
        cfg = yaml.safe_load(f)
    cfg["model"]["num_views"] = len(cfg["model"]["selected_views"])
    cfg["data"]["selected_views"] = cfg["model"]["selected_views"]
    cfg["data"]["num_views"] = cfg["model"]["num_views"]
    cfg["data"]["mask_invisible_joints"] = cfg["train"]["mask_invisible_joints"]
    return cfg


def release_config(num_views: int = 5, crop: bool = True, name: str = "ho3d", backbone: str = "resnet") -> dict:
    """In-memory equivalent of configs/release/{HO3D,DexYCB,MVHand}_HandMvNet[_HR][_wo_cam].yaml restricted to the
    keys the forward path reads (the YAML files themselves belong to the reference and are not shipped)."""
    pos_enc = ["pos2d", "crop", "sin"] if crop else ["pos2d", "sin"]
    if backbone == "hrnet":
        return {
            "name": "handmvnet",
            "data": {"name": name, "batch_size": 16, "heatmap_size": 32, "image_size": 256,
                     "selected_views": list(range(num_views)), "num_views": num_views},
            "model": {"selected_views": list(range(num_views)), "num_views": num_views, "fusion": "cross_attn",
                      "fusion_layers": 5, "pos_enc": pos_enc, "use_gcn": True, "backbone": "hrnet", "backbone_type": "w40",
                      "backbone_pretrained_path": "", "backbone_channels": [40, 80, 160, 320], "backbone_pretrained": False},
            "train": {"debug": False, "root_relative": True, "device": "cuda"},
        }
    return {
        "name": "handmvnet",
        "data": {"name": name, "batch_size": 16, "heatmap_size": 32, "image_size": 256,
                 "selected_views": list(range(num_views)), "num_views": num_views},
        "model": {"selected_views": list(range(num_views)), "num_views": num_views, "fusion": "cross_attn",
                  "fusion_layers": 5, "pos_enc": pos_enc, "use_gcn": True, "backbone": "resnet",
                  "backbone_type": "50_paper", "backbone_early_return": 3, "backbone_channels": [1024],
                  "backbone_pretrained": False},
        "train": {"debug": False, "root_relative": True, "device": "cuda"},
    }
