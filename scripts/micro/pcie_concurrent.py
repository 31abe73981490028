#!/usr/bin/env python
"""Aggregate D2H / H2D bandwidth with 1, 2, 4, 8 GPUs copying AT THE SAME TIME (one process per GPU under torchrun, pinned
host buffers, gloo barrier): separates the per-link rate from what the host side (memory write bandwidth, root complex, NUMA
placement) sustains — the limiter of the end-to-end numbers of bench.py at N > 1.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/micro/pcie_concurrent.py"""
import os, time, json
import torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("gloo")
torch.cuda.set_device(local)
n = 640 * 1024 * 1024
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h.fill_(1)
out = {}
for name, dst, src in (("D2H", h, d), ("H2D", d, h)):
    dst.copy_(src, non_blocking=True); torch.cuda.synchronize()
    for active in sorted({1, world}):  # one rank alone, then all ranks together
        if world > 1: dist.barrier()
        t0 = time.perf_counter()
        if rank < active:
            for _ in range(4): dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / 4
        t = torch.tensor([dt if rank < active else 0.0], dtype=torch.float64)
        if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[f"{name}_{active}ranks_GBs_aggregate"] = round(active * n / float(t[0]) / 1e9, 1)
if rank == 0:
    try:
        numa = open("/sys/devices/system/node/online").read().strip()
    except OSError:
        numa = "?"
    print(json.dumps({"gpus": world, "numa_nodes_online": numa, "host_threads": len(os.sched_getaffinity(0)), **out}))
if world > 1: dist.destroy_process_group()
