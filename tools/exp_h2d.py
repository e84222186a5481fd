"""Raw pinned-host -> device copy rate on this box vs chunk size (the ceiling of every e2e number)."""
import time, torch
n = 131_489_792
src = torch.empty(n, dtype=torch.uint8).pin_memory()
dst = torch.empty(n, dtype=torch.uint8, device="cuda")
for chunk_mb in (4, 16, 32, 64, 131):
    c = min(n, chunk_mb << 20)
    for _ in range(2):
        for o in range(0, n, c):
            dst[o:o + c].copy_(src[o:o + c], non_blocking=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        for o in range(0, n, c):
            dst[o:o + c].copy_(src[o:o + c], non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print(f"chunks of {chunk_mb:3d} MB: {n / dt / 1e9:.1f} GB/s ({dt * 1e3:.2f} ms per 131 MB)")
