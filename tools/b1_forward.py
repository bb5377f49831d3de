"""A handful of B=1 forwards (graphs off when HMV_NO_GRAPH=1) -- the workload tools/gpu_b1_launches.sh puts under ncu."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from handmvnet_b200 import HandMvNet  # noqa: E402
from handmvnet_b200.config import release_config  # noqa: E402

v = int(sys.argv[1]) if len(sys.argv) > 1 else 5
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6
cfg = release_config(v, True)
torch.manual_seed(0)
m = HandMvNet(cfg["train"], cfg["model"], cfg["data"], precision="bf16", micro_batch=1)
m.to("cuda:0").eval()
m.freeze()
m.prepare("cuda:0")
x = torch.randn(1, v, 3, 256, 256, device="cuda:0")
bbox = torch.tensor([220.0, 140.0, 420.0, 340.0], device="cuda:0").expand(1, v, 4).contiguous()
cam = {"intrinsic": torch.tensor([600.0, 600.0, 320.0, 240.0], device="cuda:0").expand(1, v, 4).contiguous()}
for _ in range(n):
    out = m(x, bbox, cam)
torch.cuda.synchronize()
assert torch.isfinite(out["joints_cam"]).all()
print("ok", m.launch_count() // n, "kernels per forward")
