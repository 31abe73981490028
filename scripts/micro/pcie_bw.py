"""Host<->device copy bandwidth of the box (pinned, 1-D and the 2-D form the C ABI uses), for DESIGN.md."""
import time, torch
n = 640 * 1024 * 1024
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
p = torch.empty(n, dtype=torch.uint8)
for name, dst, src in (("D2H pinned", h, d), ("H2D pinned", d, h), ("D2H pageable", p, d), ("H2D pageable", d, p)):
    dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 3
    print(f"{name}: {n / dt / 1e9:.1f} GB/s")
