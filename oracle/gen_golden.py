"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Runs only in the build container (needs /root/reference).  For every case it
  1. builds the reference HandMvNet (reference src/models/handmvnet.py:27) from the
     release YAML through `ref_shim`,
  2. loads `oracle.make_state_dict(...)` with strict=True  (proves key/shape parity
     of the 355-entry state_dict),
  3. runs `model(x, bbox, cam_params)` under forward hooks,
  4. stores the outputs plus fingerprints (moments + 256 fixed-index samples) of every
     stage tensor.
The fixtures are what pins `oracle/handmvnet_oracle.py` (tests/test_oracle_golden.py).

usage:  python oracle/gen_golden.py [case ...]
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import handmvnet_oracle as O  # noqa: E402
import ref_shim  # noqa: E402

CASES = [
    # name,            yaml,                       batch, seed_w, seed_x, randomize_norm
    ("ho3d_v5_rand",   "HO3D_HandMvNet",           1,     0,      1234,   True),
    ("ho3d_v5_plain",  "HO3D_HandMvNet",           2,     1,      77,     False),
    ("dexycb_v8_rand", "DexYCB_HandMvNet",         1,     2,      5,      True),
    ("ho3d_v5_wo_cam", "HO3D_HandMvNet_wo_cam",    1,     3,      9,      True),
    ("ho3d_v5_hr",     "HO3D_HandMvNet_HR",        1,     4,      11,     True),      # HRNet-w40, 4 feature levels, d_model 312
    ("mvhand_v4_hr_wo_cam", "MVHand_HandMvNet_HR_wo_cam", 1, 5,   13,     False),     # HRNet-w40, 4 views, no "crop"
]


def fingerprint(t: torch.Tensor, seed: int = 0, n: int = 256):
    flat = t.detach().reshape(-1).to(torch.float64)
    g = torch.Generator().manual_seed(seed)
    idx = torch.randint(0, flat.numel(), (n,), generator=g)
    return {
        "shape": np.array(t.shape, dtype=np.int64),
        "mean": np.float64(flat.mean()), "std": np.float64(flat.std()),
        "absmean": np.float64(flat.abs().mean()),
        "idx": idx.numpy(), "val": flat[idx].to(torch.float32).numpy(),
    }


def main():
    torch.set_num_threads(os.cpu_count())
    out_dir = os.path.join(os.path.dirname(HERE), "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    only = set(sys.argv[1:])                      # optional: names of the cases to (re)generate
    for name, yaml_name, batch, seed_w, seed_x, rnd in CASES:
        if only and name not in only:
            continue
        cfg = ref_shim.load_cfg(yaml_name)
        model = ref_shim.build_reference(cfg)
        ocfg = O.release_config(cfg["model"]["num_views"], "crop" in cfg["model"]["pos_enc"], cfg["model"]["backbone"])
        sd = O.make_state_dict(ocfg, seed=seed_w, randomize_norm=rnd)
        ref_keys = list(model.state_dict().keys())
        assert ref_keys == list(sd.keys()), "state_dict key order/names differ from the reference"
        model.load_state_dict(sd, strict=True)
        model.eval()
        x, bbox, intr = O.make_inputs(batch, cfg["model"]["num_views"], seed=seed_x)

        stage = {}

        def keep(key):
            def hook(_m, _inp, out):
                stage[key] = out.detach().clone()
            return hook

        hr = cfg["model"]["backbone"] == "hrnet"
        if hr:
            def keep_levels(_m, _inp, out):
                for l, t in enumerate(out):
                    stage[f"level{l}"] = t.detach().clone()
            hooks = [model.backbone.layer1.register_forward_hook(keep("hr.layer1")),
                     model.backbone.stage2.register_forward_hook(lambda _m, _i, out: stage.__setitem__("hr.stage2.0.out1", out[1].detach().clone())),
                     model.backbone.stage3.register_forward_hook(lambda _m, _i, out: stage.__setitem__("hr.stage3.3.out2", out[2].detach().clone())),
                     model.backbone.register_forward_hook(keep_levels)]
        else:
            hooks = [
                model.backbone.maxpool.register_forward_hook(keep("stem")),
                model.backbone.layer1.register_forward_hook(keep("layer1")),
                model.backbone.layer2.register_forward_hook(keep("layer2")),
                model.backbone.register_forward_hook(keep("backbone_out")),
                model.sample_nets[0].register_forward_hook(keep("sampled")),
            ]
        hooks += [
            model.pose_net.register_forward_hook(keep("heatmap")),
            model.joints_late_fusion.register_forward_pre_hook(
                lambda _m, inp: stage.__setitem__("tokens", inp[0].detach().clone())),
            model.joints_decoder.register_forward_hook(keep("joints_cam")),
        ]
        for i, layer in enumerate(model.joints_late_fusion.attn_fusion):
            hooks.append(layer.register_forward_hook(keep(f"fusion{i}")))
        with torch.no_grad():
            if "crop" in cfg["model"]["pos_enc"]:
                out = model(x, bbox, {"intrinsic": intr, "extrinsic": torch.zeros(batch, x.shape[1], 4, 4)})
            else:
                out = model(x)
        for h in hooks:
            h.remove()

        blob = {
            "meta_yaml": np.array(yaml_name), "meta_batch": np.int64(batch), "meta_seed_w": np.int64(seed_w),
            "meta_seed_x": np.int64(seed_x), "meta_randomize_norm": np.bool_(rnd),
            "meta_num_views": np.int64(cfg["model"]["num_views"]),
            "meta_crop": np.bool_("crop" in cfg["model"]["pos_enc"]),
            "meta_backbone": np.array(cfg["model"]["backbone"]),
            "meta_torch": np.array(torch.__version__),
            "out_joints_cam": out["joints_cam"].numpy(),
            "out_joints_crop_img": out["joints_crop_img"].numpy(),
            "out_heatmap_sub": out["heatmap"][..., ::4, ::4].numpy(),
            "out_heatmap_max": out["heatmap"].flatten(-2).max(-1).values.numpy(),
            "out_heatmap_argmax": out["heatmap"].flatten(-2).argmax(-1).numpy(),
        }
        for k, t in stage.items():
            for fk, fv in fingerprint(t).items():
                blob[f"stage_{k}_{fk}"] = fv
        path = os.path.join(out_dir, name + ".npz")
        np.savez_compressed(path, **blob)
        print(f"{name}: wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB); "
              f"mean|joints_cam|={out['joints_cam'].abs().mean():.4e} "
              f"mean|heatmap|={out['heatmap'].abs().mean():.4f} "
              f"backbone std={stage['level0' if hr else 'backbone_out'].std():.3f}")


if __name__ == "__main__":
    main()
