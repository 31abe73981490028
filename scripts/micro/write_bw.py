#!/usr/bin/env python
"""Write-only / read-only / copy bandwidth of the device (torch fill_, sum, copy_ on 2 GiB), for the HBM rooflines of kernels
whose traffic is all stores (CRBA: 661 MB of M per launch, 47 MB of reads)."""
import torch
n = 1 << 28  # 2 GiB of float64
a = torch.empty(n, dtype=torch.float64, device="cuda"); b = torch.empty(n, dtype=torch.float64, device="cuda")
def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
t = timeit(lambda: a.fill_(1.5)); print(f"write only  (fill_): {8*n/t/1e6:8.1f} GB/s")
t = timeit(lambda: a.zero_()); print(f"write only  (zero_): {8*n/t/1e6:8.1f} GB/s")
t = timeit(lambda: a.sum()); print(f"read only   (sum):   {8*n/t/1e6:8.1f} GB/s")
t = timeit(lambda: b.copy_(a)); print(f"copy (read + write): {16*n/t/1e6:8.1f} GB/s")
# 661 MB, the size of one CRBA output at 65 536 x simple_humanoid
m = 65536 * 1225
c = torch.empty(m, dtype=torch.float64, device="cuda")
t = timeit(lambda: c.fill_(0.25), 20); print(f"write only, 642 MB:  {8*m/t/1e6:8.1f} GB/s ({t:.4f} ms)")
