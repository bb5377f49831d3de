"""Fusion-transformer micro-benchmark (BASELINE.json config 5): CrossAttentionFusion alone (reference fusion.py:7-30),
views 2..16 -> 42..336 tokens per sample, through the stage API (hmv_stage_run FUSION).

    python tools/bench_fusion.py [batch] [iters]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from handmvnet_b200 import HandMvNet  # noqa: E402
from handmvnet_b200.config import release_config  # noqa: E402


def fusion_flops(v, d=524, inner=1024, hidden=128, layers=5):
    """2*MAC of one sample: projections, attention core, out-proj, feed-forward (reference layers.py:202-237)."""
    s = 21 * v
    half = (layers - 1) // 2
    total = 0
    for i in range(layers):
        if i < half:
            nq, nk = s, s
        elif i == half:
            nq, nk = 21, s - 21
        else:
            nq, nk = 21, 21
        proj = 2 * d * inner * (nq + 2 * nk)
        att = 2 * 2 * nq * nk * inner
        outp = 2 * nq * inner * d
        ff = 2 * 2 * nq * d * hidden
        total += proj + att + outp + ff
    return total


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    rows = []
    for v in (2, 4, 5, 8, 16):
        cfg = release_config(v, True)
        torch.manual_seed(0)
        m = HandMvNet(cfg["train"], cfg["model"], cfg["data"], precision="bf16", micro_batch=batch)
        m.to("cuda:0").eval()
        m.freeze()
        m.prepare("cuda:0")
        tok = torch.randn(batch, 21 * v, m.feat_dim, device="cuda:0")
        m.tensor_set("tokens", tok, batch)
        for _ in range(5):
            m.stage_run("fusion", batch)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            m.stage_run("fusion", batch)
        e1.record()
        e1.synchronize()
        m.synchronize()
        ms = e0.elapsed_time(e1) / iters
        fl = fusion_flops(v, d=m.feat_dim) * batch
        rows.append({"views": v, "tokens": 21 * v, "batch": batch, "ms": ms, "samples_per_s": batch / ms * 1e3,
                     "gflop_per_sample": fl / batch * 1e-9, "tflops": fl / ms * 1e-9})
        print(json.dumps(rows[-1]), flush=True)
        del m
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
