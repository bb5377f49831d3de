"""Experiment: two handles, each on its own stream with half of the SMs (HMV_NUM_SMS), each taking half of the batch,
so that HBM-bound kernels of one half overlap tensor-bound kernels of the other.  Prints poses/s for both layouts."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from handmvnet_b200 import HandMvNet  # noqa: E402
from handmvnet_b200.config import release_config  # noqa: E402


def build(mb):
    cfg = release_config(5, True)
    torch.manual_seed(0)
    m = HandMvNet(cfg["train"], cfg["model"], cfg["data"], precision="bf16", micro_batch=mb)
    m.to("cuda").eval()
    m.freeze()
    m.prepare(torch.device("cuda"))
    return m


def run(models, streams, xs, bb, it, steps):
    for _ in range(3):
        for m, s, x in zip(models, streams, xs):
            with torch.cuda.stream(s):
                m(x, bb[: x.shape[0]], {"intrinsic": it[: x.shape[0]]})
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        for m, s, x in zip(models, streams, xs):
            with torch.cuda.stream(s):
                m(x, bb[: x.shape[0]], {"intrinsic": it[: x.shape[0]]})
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / steps * 1e3


def main():
    B = 64
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(B, 5, 3, 256, 256, device="cuda", generator=g)
    bb = torch.tensor([100.0, 100.0, 300.0, 300.0], device="cuda").repeat(B, 5, 1)
    it = torch.tensor([600.0, 600.0, 320.0, 240.0], device="cuda").repeat(B, 5, 1)
    sms = int(sys.argv[1]) if len(sys.argv) > 1 else 74
    os.environ.pop("HMV_NUM_SMS", None)
    single = build(64)
    t = run([single], [torch.cuda.Stream()], [x], bb, it, 20)
    print(f"single handle, 148 SMs: {t:.2f} ms/step  {B / t * 1e3:.0f} poses/s")
    del single
    os.environ["HMV_NUM_SMS"] = str(sms)
    a, b = build(32), build(32)
    t = run([a, b], [torch.cuda.Stream(), torch.cuda.Stream()], [x[:32], x[32:]], bb, it, 20)
    print(f"two handles x {sms} SMs:  {t:.2f} ms/step  {B / t * 1e3:.0f} poses/s")


if __name__ == "__main__":
    main()
