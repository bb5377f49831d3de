"""Inference-speed driver with the reference's CLI (`--config <yaml>`), mirroring src/eval_fps.py:21-106.

Differences, all deliberate: `n_views` comes from the config (the reference hard-codes 8 and crashes
a 5-view model), bbox / intrinsics are valid synthetic values instead of randn / uninitialised memory, the MANO
mesh post-processing is not part of the timed region, and timing uses CUDA events.

    python -m handmvnet_b200.eval_fps --config configs/release/HO3D_HandMvNet.yaml [--batch 1] [--precision bf16]
"""
import argparse

import torch

from . import HandMvNet, load_config


def main():
    ap = argparse.ArgumentParser(description="Configuration args.")
    ap.add_argument("--config", type=str, required=True, help="Path to the YAML configuration file")
    ap.add_argument("--num-gpus", type=int, default=1, help="Number of GPUs")
    ap.add_argument("--checkpoint", type=str, help="Path to the model checkpoint")
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--runs", type=int, default=1000)
    args = ap.parse_args()
    cfg = load_config(args.config)
    cfg["model"]["backbone_pretrained"] = False
    cfg["train"]["device"] = "cuda"
    model = HandMvNet(cfg["train"], cfg["model"], cfg["data"], precision=args.precision, micro_batch=min(args.batch, 16))
    if args.checkpoint:
        from .checkpoint import load_checkpoint_with_legacy_fix
        load_checkpoint_with_legacy_fix(args.checkpoint, model, "cpu")
    model.to("cuda").eval()
    model.freeze()
    n_views, size = cfg["model"]["num_views"], cfg["data"]["image_size"]
    x = torch.randn(args.batch, n_views, 3, size, size, device="cuda")
    bbox = torch.tensor([220.0, 140.0, 420.0, 340.0], device="cuda").expand(args.batch, n_views, 4).contiguous()
    cam = {"intrinsic": torch.tensor([600.0, 600.0, 320.0, 240.0], device="cuda").expand(args.batch, n_views, 4).contiguous()}
    with torch.no_grad():
        for _ in range(args.warmup):
            out = model(x, bbox, cam)
        torch.cuda.synchronize()
        times = []
        for _ in range(args.runs):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = model(x, bbox, cam)
            e1.record()
            e1.synchronize()
            times.append(e0.elapsed_time(e1))
    times.sort()
    print("-------------------------------------------------")
    print(f"Batch size: {x.shape[0]}")
    print(f"Camera views: {n_views}")
    print(f"Average FPS: {1000.0 * args.batch * len(times) / sum(times):.3f}")
    print(f"p50 latency: {times[len(times) // 2]:.3f} ms   p99: {times[int(len(times) * 0.99)]:.3f} ms")
    print(f"joints_cam: {tuple(out['joints_cam'].shape)}")
    print("-------------------------------------------------")


if __name__ == "__main__":
    main()
