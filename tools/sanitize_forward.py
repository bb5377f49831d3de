"""One small forward of every kernel on the path, meant to run under compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool memcheck python tools/sanitize_forward.py [bf16|fp32] [batch]

micro_batch = batch = 2 by default, 5 views, both outputs checked finite.  HMV_NO_GRAPH=1 is set so that every kernel is
launched eagerly (graph capture under the sanitizer adds nothing)."""
import os
import sys

os.environ.setdefault("HMV_NO_GRAPH", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from handmvnet_b200 import HandMvNet  # noqa: E402
from handmvnet_b200.config import release_config  # noqa: E402


def main():
    precision = sys.argv[1] if len(sys.argv) > 1 else "bf16"
    b = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    cfg = release_config(5, True)
    torch.manual_seed(0)
    m = HandMvNet(cfg["train"], cfg["model"], cfg["data"], precision=precision, micro_batch=b)
    m.to("cuda:0").eval()
    m.freeze()
    m.prepare("cuda:0")
    x = torch.randn(b, 5, 3, 256, 256, device="cuda:0")
    bbox = torch.tensor([220.0, 140.0, 420.0, 340.0], device="cuda:0").expand(b, 5, 4).contiguous()
    cam = {"intrinsic": torch.tensor([600.0, 600.0, 320.0, 240.0], device="cuda:0").expand(b, 5, 4).contiguous()}
    out = m(x, bbox, cam)
    m.synchronize()
    assert torch.isfinite(out["joints_cam"]).all() and torch.isfinite(out["heatmap"]).all()
    xu = torch.randint(0, 256, (b, 5, 3, 256, 256), dtype=torch.uint8, device="cuda:0")
    out = m(xu, bbox, cam)
    m.synchronize()
    assert torch.isfinite(out["joints_cam"]).all()
    print(f"sanitize_forward[{precision}] B={b}: ok, {m.launch_count()} kernels")


if __name__ == "__main__":
    main()
