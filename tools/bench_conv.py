"""Per-layer micro-benchmark through hmv_conv_bn_act (the model's own conv kernels): time, TFLOP/s and the
algorithmic HBM GB/s (input + output + residual, bf16) of each backbone geometry class.

    python tools/bench_conv.py [n_img] [case-substring]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
from gpu_util import conv_bn_act  # noqa: E402

CASES = [
    # name,              cin, cout, k, s, H,  W,  res
    ("l1.conv1",          256, 64, 1, 1, 64, 64, False),
    ("l1.conv2",          64, 64, 3, 1, 64, 64, False),
    ("l1.conv3",          64, 256, 1, 1, 64, 64, True),
    ("l1.down",           64, 256, 1, 1, 64, 64, False),
    ("l2.0.conv1",        256, 128, 1, 1, 64, 64, False),
    ("l2.0.conv2s2",      128, 128, 3, 2, 64, 64, False),
    ("l2.0.down_s2",      256, 512, 1, 2, 64, 64, False),
    ("l2.conv1",          512, 128, 1, 1, 32, 32, False),
    ("l2.conv2",          128, 128, 3, 1, 32, 32, False),
    ("l2.conv3",          128, 512, 1, 1, 32, 32, True),
    ("l3.0.conv1",        512, 256, 1, 1, 32, 32, False),
    ("l3.0.down",         512, 1024, 1, 1, 32, 32, False),
    ("l3.conv1",          1024, 256, 1, 1, 32, 32, False),
    ("l3.conv2",          256, 256, 3, 1, 32, 32, False),
    ("l3.conv3",          256, 1024, 1, 1, 32, 32, True),
    ("pose0",             1024, 512, 1, 1, 32, 32, False),
]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 160
    flt = sys.argv[2] if len(sys.argv) > 2 else ""
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    print(f"{'layer':14s} {'ms':>8s} {'TFLOP/s':>9s} {'GB/s':>8s}   (n_img={n})")
    for name, cin, cout, k, s, h, w, res in CASES:
        if flt and flt not in name:
            continue
        g = torch.Generator(device="cuda").manual_seed(0)
        x = torch.randn(n, cin, h, w, device="cuda", generator=g)
        wt = torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5
        r = torch.randn(n, cout, h // s, w // s, device="cuda", generator=g) if res else None
        _, ms = conv_bn_act("bf16", x, wt, None, None, r, stride=s, relu=True, iters=iters)
        m = n * (h // s) * (w // s)
        flop = 2.0 * m * cout * cin * k * k
        byt = 2.0 * (n * h * w * cin + m * cout * (2 if res else 1))
        print(f"{name:14s} {ms:8.4f} {flop / ms * 1e-9:9.1f} {byt / ms * 1e-6:8.0f}")
        del x, wt, r
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
