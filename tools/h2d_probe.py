"""Pinned host->device copy bandwidth on this box (sizes of one bench step and of its chunks)."""
import torch
for mb in (4, 16, 63, 126, 252):
    n = mb * 1024 * 1024 // 4
    h = torch.empty(n, dtype=torch.float32).pin_memory()
    d = torch.empty(n, dtype=torch.float32, device="cuda")
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"H2D {mb} MiB: {ms:.3f} ms  {mb * 1.048576 / ms:.1f} GB/s")
