"""p50 / p99 latency of one B=1 forward (BASELINE.json metric 'p50 B=1 latency'), CUDA-event timed, device-resident
inputs, for the 5-view HO3D and the 8-view DexYCB release configurations.

    python tools/bench_latency.py [iters]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from handmvnet_b200 import HandMvNet  # noqa: E402
from handmvnet_b200.config import release_config  # noqa: E402


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    for v, b in ((5, 1), (8, 1), (5, 8)):
        cfg = release_config(v, True)
        torch.manual_seed(0)
        m = HandMvNet(cfg["train"], cfg["model"], cfg["data"], precision="bf16", micro_batch=b)
        m.to("cuda:0").eval()
        m.freeze()
        m.prepare("cuda:0")
        x = torch.randn(b, v, 3, 256, 256, device="cuda:0")
        bbox = torch.tensor([220.0, 140.0, 420.0, 340.0], device="cuda:0").expand(b, v, 4).contiguous()
        cam = {"intrinsic": torch.tensor([600.0, 600.0, 320.0, 240.0], device="cuda:0").expand(b, v, 4).contiguous()}
        for _ in range(50):
            m(x, bbox, cam)
        torch.cuda.synchronize()
        times = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = m(x, bbox, cam)
            e1.record()
            e1.synchronize()
            times.append(e0.elapsed_time(e1))
        times.sort()
        print(json.dumps({"views": v, "batch": b, "p50_ms": times[len(times) // 2], "p99_ms": times[int(len(times) * 0.99)],
                          "min_ms": times[0], "poses_per_s_at_p50": b / times[len(times) // 2] * 1e3,
                          "launches_per_forward": m.launch_count() // (iters + 50)}), flush=True)
        del m
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
