"""Import shim for the upstream reference (test infrastructure only).

Only usable where /root/reference exists (the build container).  It stubs the
third-party modules the reference imports at module scope but which are absent
from this image (lightning, manopth, matplotlib, plotly, transforms3d) so that
`models.handmvnet.HandMvNet` (reference src/models/handmvnet.py:27) can be
constructed and run on CPU.  Nothing on the product path imports this file.
"""
import os
import sys
import types

import torch
import yaml

REF_ROOT = os.environ.get("HMV_REFERENCE_ROOT", "/root/reference")


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install():
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")

    class LightningModule(torch.nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, *a, **k):
            pass

        def freeze(self):
            for p in self.parameters():
                p.requires_grad = False
            self.eval()

    class LightningDataModule:
        pass

    if "lightning" not in sys.modules:
        L = _stub("lightning", LightningModule=LightningModule, LightningDataModule=LightningDataModule)
        _stub("lightning.pytorch")
        _stub("lightning.pytorch.utilities")
        _stub("lightning.pytorch.utilities.model_summary", ModelSummary=object)
        L.pytorch = sys.modules["lightning.pytorch"]
    for name in ["manopth", "plotly", "matplotlib", "transforms3d", "matplotlib.pyplot",
                 "plotly.graph_objs", "plotly.graph_objects", "matplotlib.patches",
                 "mpl_toolkits", "mpl_toolkits.mplot3d", "transforms3d.axangles", "transforms3d.euler",
                 "transforms3d.quaternions"]:
        if name not in sys.modules:
            _stub(name)
    if not hasattr(sys.modules["manopth"], "manolayer"):
        ml = _stub("manopth.manolayer", ManoLayer=object)
        sys.modules["manopth"].manolayer = ml
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    src = os.path.join(REF_ROOT, "src")
    if src not in sys.path:
        sys.path.insert(0, src)


def load_cfg(name="HO3D_HandMvNet"):
    """YAML -> dict exactly as reference src/config.py:35-51 derives it."""
    with open(os.path.join(REF_ROOT, "configs", "release", name + ".yaml")) as f:
        cfg = yaml.safe_load(f)
    cfg["model"]["num_views"] = len(cfg["model"]["selected_views"])
    cfg["data"]["selected_views"] = cfg["model"]["selected_views"]
    cfg["data"]["num_views"] = cfg["model"]["num_views"]
    cfg["data"]["mask_invisible_joints"] = cfg["train"]["mask_invisible_joints"]
    cfg["model"]["backbone_pretrained"] = False  # no network
    cfg["train"]["device"] = "cpu"
    return cfg


def build_reference(cfg):
    install()
    from models.handmvnet import HandMvNet
    m = HandMvNet(cfg["train"], cfg["model"], cfg["data"])
    m.eval()
    return m
