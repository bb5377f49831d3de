"""Host-side (no GPU) checks of the drop-in boundary: state_dict contract, constructor error conventions of
reference src/models/handmvnet.py:28-125, no CPU fallback, and that the C-ABI library exports every symbol
include/handmvnet_b200.h declares."""
import ctypes
import os
import re

import pytest
import torch

import handmvnet_oracle as O
from handmvnet_b200 import HandMvNet, _lib
from handmvnet_b200.config import release_config

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _model(views=5, crop=True, **kw):
    cfg = release_config(views, crop)
    return HandMvNet(cfg["train"], cfg["model"], cfg["data"], **kw), cfg


@pytest.mark.parametrize("views,crop", [(5, True), (8, True), (5, False)])
def test_state_dict_contract(views, crop):
    m, _ = _model(views, crop)
    ocfg = O.release_config(views, crop)
    spec = {k: tuple(s) for k, s, _ in O._state_dict_spec(ocfg)}
    sd = m.state_dict()
    assert len(sd) == 355 and set(sd.keys()) == set(spec.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == spec[k], k
    assert m.feat_dim == (524 if crop else 514)
    # a reference-shaped state_dict loads strictly (src/eval.py:46-50)
    m.load_state_dict(O.make_state_dict(ocfg, seed=3), strict=True)
    assert "joints_late_fusion.pos_encoding.pe" not in sd           # plain attribute, like the reference
    assert torch.equal(m.joints_late_fusion.pos_encoding.pe[0], O.positional_table(m.feat_dim, 21 * views))


def test_constructor_error_conventions():
    cfg = release_config(5, True)
    bad = dict(cfg["model"], backbone="vgg")
    with pytest.raises(AssertionError):
        HandMvNet(cfg["train"], bad, cfg["data"])
    with pytest.raises(AssertionError):
        HandMvNet(cfg["train"], dict(cfg["model"], backbone_type="101"), cfg["data"])
    with pytest.raises(NotImplementedError):
        HandMvNet(cfg["train"], dict(cfg["model"], fusion="concat"), cfg["data"])
    with pytest.raises(NotImplementedError):
        HandMvNet(cfg["train"], cfg["model"], dict(cfg["data"], name="freihand"))
    with pytest.raises(AssertionError):            # fusion.py:11
        HandMvNet(cfg["train"], dict(cfg["model"], fusion_layers=4), cfg["data"])
    with pytest.raises(NotImplementedError):       # scoped out
        HandMvNet(dict(cfg["train"], root_relative=False), cfg["model"], cfg["data"])
    with pytest.raises(ValueError):                # HRNet widths must match the type (w40: 40/80/160/320)
        HandMvNet(cfg["train"], dict(cfg["model"], backbone="hrnet", backbone_type="w40", backbone_channels=[1024]), cfg["data"])
    with pytest.raises(Exception, match="HRNet only supports"):    # backbones/hrnet.py:444
        HandMvNet(cfg["train"], dict(cfg["model"], backbone="hrnet", backbone_type="w18", backbone_channels=[18, 36, 72, 144]), cfg["data"])


@pytest.mark.parametrize("views,crop", [(5, True), (4, False)])
def test_state_dict_contract_hrnet(views, crop):
    """The `*_HR*` release configs: 1941 keys in the reference's order, strict load of a reference-shaped state_dict."""
    cfg = release_config(views, crop, backbone="hrnet")
    m = HandMvNet(cfg["train"], cfg["model"], cfg["data"])
    ocfg = O.release_config(views, crop, "hrnet")
    spec = [(k, tuple(s)) for k, s, _ in O._state_dict_spec(ocfg)]
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == spec and len(spec) == 1941
    assert m.feat_dim == (312 if crop else 302)
    m.load_state_dict(O.make_state_dict(ocfg, seed=3), strict=True)


def test_no_cpu_fallback_and_input_validation():
    m, _ = _model()
    m.eval()
    x = torch.zeros(1, 5, 3, 256, 256)
    bbox = torch.zeros(1, 5, 4)
    cam = {"intrinsic": torch.ones(1, 5, 4)}
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(x, bbox, cam)
    with pytest.raises(ValueError, match="views"):
        m(torch.zeros(1, 8, 3, 256, 256), torch.zeros(1, 8, 4), {"intrinsic": torch.ones(1, 8, 4)})
    with pytest.raises(ValueError, match="crop"):
        m(x)                                         # 'crop' positional encoding needs bbox + intrinsics
    with pytest.raises(RuntimeError):                # parameter holders never compute
        m.backbone(torch.zeros(1, 3, 256, 256))
    m.freeze()
    assert not any(p.requires_grad for p in m.parameters()) and not m.training


def test_library_exports_every_header_symbol():
    header = open(os.path.join(ROOT, "include", "handmvnet_b200.h")).read()
    declared = sorted(set(re.findall(r"\b(hmv_[a-z0-9_]+)\s*\(", header)))
    assert declared == sorted(_lib.EXPORTS)
    lib = _lib.load()                                # raises if the .so was not built
    for name in declared:
        assert isinstance(getattr(lib, name), ctypes._CFuncPtr), name
    assert b"sm_100a" in lib.hmv_version()


def test_create_without_gpu_reports_error():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = _lib.load()
    cfg = _lib.HmvConfig(num_views=5, image_size=256, heatmap_size=32, use_pos2d=1, use_crop=1, use_sin=1,
                         fusion_layers=5, precision=0, micro_batch=1, device=0)
    h = ctypes.c_void_p()
    assert lib.hmv_create(ctypes.byref(cfg), ctypes.byref(h)) != 0
    assert len(lib.hmv_last_error()) > 0
    with pytest.raises(RuntimeError):
        _lib.check(1, "hmv_create")


def test_checkpoint_loader_mirrors_the_reference_helper(tmp_path):
    """src/eval.py:27-52: `load_checkpoint_with_legacy_fix(checkpoint_path, model, device)` loads a Lightning checkpoint
    strictly and falls back to the legacy key remap (`pose_net.conv.` -> `pose_net.`, `sample_net.` -> `sample_nets.0.`)."""
    from handmvnet_b200.checkpoint import load_checkpoint_with_legacy_fix
    m, _ = _model()
    sd = O.make_state_dict(O.release_config(5, True), seed=7)
    legacy = {}
    for k, v in sd.items():
        k = k.replace("pose_net.", "pose_net.conv.") if k.startswith("pose_net.") else k
        k = "sample_net." + k[len("sample_nets.0."):] if k.startswith("sample_nets.0.") else k
        legacy[k] = v
    for name, blob in (("new.ckpt", {"state_dict": dict(sd)}), ("legacy.ckpt", {"state_dict": legacy})):
        path = tmp_path / name
        torch.save(blob, path)
        load_checkpoint_with_legacy_fix(str(path), m, "cpu")
        assert torch.equal(m.state_dict()["pose_net.0.weight"], sd["pose_net.0.weight"])
        assert torch.equal(m.state_dict()["sample_nets.0.conv.0.weight"], sd["sample_nets.0.conv.0.weight"])


def test_every_runtime_switch_of_the_library_is_documented():
    """Every `getenv("HMV_*")` of the CUDA library appears in tools/README.md (the one index of the switches), so an A/B
    knob cannot ship undocumented; and the index names no switch the library no longer reads."""
    import glob
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    read = set()
    for path in glob.glob(os.path.join(root, "handmvnet_b200", "csrc", "*.cu*")):
        read |= set(re.findall(r'getenv\("(HMV_[A-Z0-9_]+)"\)', open(path).read()))
    assert len(read) >= 15
    index = open(os.path.join(root, "tools", "README.md")).read()
    switches = index[index.index("Environment switches of the library"):]
    undocumented = sorted(s for s in read if s not in switches)
    assert not undocumented, f"switches missing from tools/README.md: {undocumented}"
    host_side = {"HMV_LIB_PATH", "HMV_BENCH_WATCHDOG"}        # read by handmvnet_b200/_lib.py and bench.py
    stale = sorted(s for s in set(re.findall(r"`(HMV_[A-Z0-9_]+)", switches)) if s not in read | host_side)
    assert not stale, f"tools/README.md lists switches the library does not read: {stale}"


REF_CONFIGS = "/root/reference/configs/release"


@pytest.mark.skipif(not os.path.isdir(REF_CONFIGS), reason="the reference checkout is only present in the build container")
def test_all_release_yaml_files_construct_the_drop_in():
    """`load_config` + the constructor accept the reference's 12 release YAML files unchanged (reference src/config.py:35-51,
    handmvnet.py:28-156): view count, token width and state_dict size per family, and the in-memory `release_config()`
    used by the GPU tests / bench agrees with the YAML on every key the forward path reads."""
    import glob
    from handmvnet_b200 import load_config
    paths = sorted(glob.glob(os.path.join(REF_CONFIGS, "*.yaml")))
    assert len(paths) == 12
    views = {"HO3D": 5, "DexYCB": 8, "MVHand": 4}
    for path in paths:
        name = os.path.basename(path)
        cfg = load_config(path)
        hr, crop = "_HR" in name, "wo_cam" not in name
        assert cfg["model"]["num_views"] == views[name.split("_")[0]] == cfg["data"]["num_views"], name
        m = HandMvNet(cfg["train"], cfg["model"], cfg["data"])
        assert len(m.state_dict()) == (1941 if hr else 355), name
        assert m.feat_dim == ((312 if crop else 302) if hr else (524 if crop else 514)), name
        mem = release_config(cfg["model"]["num_views"], crop, backbone="hrnet" if hr else "resnet")
        for key in ("fusion", "fusion_layers", "pos_enc", "use_gcn", "backbone", "backbone_type", "backbone_channels", "num_views"):
            assert mem["model"][key] == cfg["model"][key], (name, key)
        for key in ("heatmap_size", "image_size"):
            assert mem["data"][key] == cfg["data"][key], (name, key)
        assert mem["train"]["root_relative"] == cfg["train"]["root_relative"], name
