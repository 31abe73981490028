#!/usr/bin/env python
"""Where the kernels generated for a model start to beat the small-batch paths (cooperative kernels / generic thread kernels):
device-resident time per call against the batch size, specialised pool with min_batch = 1 against an unspecialised pool.
    python scripts/gen_small_batch.py [models...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import pinocchio_b200 as pb
from conftest import load_model, random_inputs
for name in (sys.argv[1:] or ["simple_humanoid_ff", "humanoid_random", "manipulator"]):
    model = load_model(name)
    pools = {"generic": pb.ModelPool(model, [0]), "generated": pb.ModelPool(model, [0])}
    pools["generated"].specialize(["rnea", "aba", "crba"], min_batch=1)
    for p in pools.values(): p.set_stream(torch.cuda.current_stream().cuda_stream)
    for B in (128, 256, 512, 1024, 2048, 4096, 8192, 16384):
        q, v, x = random_inputs(model, B, 1)
        tq, tv, tx = (torch.from_numpy(np.ascontiguousarray(t.T)).cuda() for t in (q, v, x))
        line = f"{name} B={B:6d}:"
        for algo in ("rnea", "aba", "crba"):
            for mode, pool in pools.items():
                fn = {"rnea": lambda o: pb.rneaInParallel(1, pool, tq, tv, tx, o, async_=True),
                      "aba": lambda o: pb.abaInParallel(1, pool, tq, tv, tx, o, async_=True),
                      "crba": lambda o: pb.crbaInParallel(1, pool, tq, o, async_=True)}[algo]
                out = torch.empty((B, model.nv * model.nv if algo == "crba" else model.nv), dtype=torch.float64, device="cuda")
                for _ in range(5): fn(out)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(50): fn(out)
                e1.record(); torch.cuda.synchronize()
                line += f" {algo}[{mode}] {e0.elapsed_time(e1) / 50 * 1e3:6.1f} us |"
        print(line, flush=True)
    for p in pools.values(): p.close()
