"""Measure pinned host->device and device->host copy bandwidth on the box (floor of the e2e number)."""
import torch, time
for mb in (64, 256, 1024, 4096):
    n = mb * 1024 * 1024
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(2):
        d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    h2d = 3 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9
    e0.record()
    for _ in range(3):
        h.copy_(d, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    d2h = 3 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9
    print(f"{mb} MiB: H2D {h2d:.1f} GB/s  D2H {d2h:.1f} GB/s", flush=True)
    del h, d
