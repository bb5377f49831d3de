"""Bring-up diagnostics (run on the GPU box): exercises every kernel class and prints error patterns instead of
asserting, so one gpurun call localises a fault.  Writes gpurun_out/diag.txt."""
import os
import sys
import time
import traceback

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import handmvnet_oracle as O  # noqa: E402
from gpu_util import build_pair, conv_bn_act, rel_l2  # noqa: E402

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
LOG = open(os.path.join(ROOT, "gpurun_out", "diag.txt"), "w")


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    LOG.write(s + "\n")
    LOG.flush()


def conv_diag(precision, cin, cout, k, stride, h, w, n=2):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, cin, h, w, generator=g).bfloat16().float()
    wt = (torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5).bfloat16().float()
    ref = F.conv2d(x, wt, stride=stride, padding=k // 2)
    try:
        t0 = time.time()
        out, ms = conv_bn_act(precision, x.cuda(), wt.cuda(), None, None, None, stride=stride, relu=False, iters=3)
        out = out.cpu()
        err = rel_l2(out, ref)
        log(f"conv[{precision}] cin={cin} cout={cout} k={k} s={stride} {h}x{w}: rel-L2 {err:.3e}  ({ms*1e3:.1f} us/launch, wall {time.time()-t0:.1f}s)")
        if err > 1e-2:
            d = (out - ref)
            # error by output channel block of 16, by pixel row, by image
            cb = d.pow(2).sum(dim=(0, 2, 3)).reshape(-1, min(16, cout)).sum(1) / (ref.pow(2).sum(dim=(0, 2, 3)).reshape(-1, min(16, cout)).sum(1) + 1e-20)
            log("   rel err^2 per 16-channel block:", [f"{v:.2e}" for v in cb.tolist()[:16]])
            pr = d.pow(2).sum(dim=(0, 1, 3)) / (ref.pow(2).sum(dim=(0, 1, 3)) + 1e-20)
            log("   rel err^2 per output row:", [f"{v:.1e}" for v in pr.tolist()[:16]])
            pc = d.pow(2).sum(dim=(0, 1, 2)) / (ref.pow(2).sum(dim=(0, 1, 2)) + 1e-20)
            log("   rel err^2 per output col:", [f"{v:.1e}" for v in pc.tolist()[:16]])
            log("   out[0,0,0,:8] =", out[0, 0, 0, :8].tolist())
            log("   ref[0,0,0,:8] =", ref[0, 0, 0, :8].tolist())
            log("   out[0,:8,0,0] =", out[0, :8, 0, 0].tolist())
            log("   ref[0,:8,0,0] =", ref[0, :8, 0, 0].tolist())
            log("   ratio mean |out|/|ref| =", float(out.abs().mean() / ref.abs().mean()))
    except Exception as e:  # noqa: BLE001
        log(f"conv[{precision}] cin={cin} cout={cout} k={k} s={stride} {h}x{w}: EXCEPTION {e}")


def backbone_bisect(precision):
    try:
        m, ocfg, sd = build_pair(5, True, precision, micro_batch=1, seed=0)
    except Exception as e:  # noqa: BLE001
        log(f"[{precision}] build failed: {e}")
        return None
    x, bbox, intr = O.make_inputs(1, 5, seed=1234)
    ximg = x.reshape(-1, 3, 256, 256)
    taps = {}
    O.backbone(sd, ximg, taps, per_layer=True)
    names = m.debug_backbone_steps()
    for i, nm in enumerate(names):
        if nm == "pack_input":
            continue
        try:
            out = m.debug_backbone(ximg.cuda(), i + 1).cpu()
            m.synchronize()
            ref = taps[nm]
            log(f"  [{precision}] step {i:2d} {nm:22s} shape {tuple(out.shape)} rel-L2 {rel_l2(out, ref):.3e}  |ref| {float(ref.abs().mean()):.3e}")
        except Exception as e:  # noqa: BLE001
            log(f"  [{precision}] step {i:2d} {nm}: EXCEPTION {e}")
            break
    return m, ocfg, sd, (x, bbox, intr)


def main():
    phase = sys.argv[1] if len(sys.argv) > 1 else "all"
    global LOG
    LOG.close()
    LOG = open(os.path.join(ROOT, "gpurun_out", f"diag_{phase}.txt"), "w")
    log("device:", torch.cuda.get_device_name(0), "phase:", phase)
    cases = [(64, 64, 1, 1, 64, 64), (128, 128, 1, 1, 32, 32), (1024, 256, 1, 1, 32, 32), (256, 1024, 1, 1, 32, 32),
             (512, 21, 1, 1, 32, 32), (64, 64, 3, 1, 64, 64), (128, 128, 3, 1, 32, 32), (256, 256, 3, 1, 32, 32),
             (128, 128, 3, 2, 64, 64), (256, 512, 1, 2, 64, 64)]
    for prec in ("fp32", "bf16"):
        if phase in ("all", "conv_" + prec):
            for case in cases:
                conv_diag(prec, *case)
    for prec in ("fp32", "bf16"):
        if phase not in ("all", "model_" + prec):
            continue
        log(f"--- backbone bisect {prec} ---")
        r = backbone_bisect(prec)
        if r is None:
            continue
        m, ocfg, sd, (x, bbox, intr) = r
        try:
            ref, taps = O.forward(sd, ocfg, x, bbox, intr, return_taps=True)
            out = m(x.cuda(), bbox.cuda(), {"intrinsic": intr.cuda()})
            m.synchronize()
            log(f"  [{prec}] e2e heatmap rel-L2 {rel_l2(out['heatmap'], ref['heatmap']):.3e}")
            log(f"  [{prec}] e2e feat rel-L2 {rel_l2(m.tensor_get('feat', 1), taps['backbone_out']):.3e}")
            log(f"  [{prec}] e2e xy max abs {float((out['joints_crop_img'].cpu() - ref['joints_crop_img']).abs().max()):.3e}")
            log(f"  [{prec}] e2e tokens rel-L2 {rel_l2(m.tensor_get('tokens', 1), taps['tokens_pe']):.3e}")
            log(f"  [{prec}] e2e fused rel-L2 {rel_l2(m.tensor_get('fused', 1), taps['fused']):.3e}")
            log(f"  [{prec}] e2e joints max abs mm {float((out['joints_cam'].cpu() - ref['joints_cam']).abs().max()) * 1e3:.4f}  rel {rel_l2(out['joints_cam'], ref['joints_cam']):.3e}")
        except Exception:  # noqa: BLE001
            log(traceback.format_exc())


if __name__ == "__main__":
    main()
