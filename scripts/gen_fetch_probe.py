#!/usr/bin/env python
"""Is the generated ABA kernel's instruction supply limited per SM or by all SMs asking L2 for the same lines at once?
Times one round of the persistent grid at B = 256 * k for k = 1, 2, 4, ..., 148 CTAs (one CTA per SM, 256 configurations each)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import pinocchio_b200 as pb
from conftest import load_model, random_inputs
model = load_model(sys.argv[1] if len(sys.argv) > 1 else "simple_humanoid_ff")
pool = pb.ModelPool(model, [0]); pool.specialize(["rnea", "aba"], min_batch=1)
pool.set_stream(torch.cuda.current_stream().cuda_stream)
for k in (1, 2, 4, 8, 16, 32, 64, 100, 148, 296):
    B = 256 * k
    q, v, x = random_inputs(model, B, 1)
    tq, tv, tx = (torch.from_numpy(np.ascontiguousarray(t.T)).cuda() for t in (q, v, x))
    for algo, fn in (("rnea", pb.rneaInParallel), ("aba", pb.abaInParallel)):
        out = fn(1, pool, tq, tv, tx)
        for _ in range(3): fn(1, pool, tq, tv, tx, out, async_=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn(1, pool, tq, tv, tx, out, async_=True)
        e1.record(); torch.cuda.synchronize()
        print(f"{algo} CTAs={k:4d} B={B:6d}: {e0.elapsed_time(e1)/20*1e3:8.1f} us", flush=True)
