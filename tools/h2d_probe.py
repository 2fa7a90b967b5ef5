"""Raw pinned host->device bandwidth of this box (one big copy, and 1.6 MB copies back to back) -- the ceiling of e2e."""
import time
import torch
x = torch.empty(2 << 30, dtype=torch.uint8).pin_memory()
y = torch.empty(2 << 30, dtype=torch.uint8, device="cuda")
for name, piece in (("one 2 GiB copy", 2 << 30), ("1.6 MB copies", 1_600_000), ("16 MB copies", 16 << 20)):
    torch.cuda.synchronize()
    for rep in range(2):
        t0 = time.perf_counter()
        off = 0
        while off + piece <= x.numel():
            y[off: off + piece].copy_(x[off: off + piece], non_blocking=True)
            off += piece
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    print(f"{name}: {off / dt / 1e9:.1f} GB/s")
