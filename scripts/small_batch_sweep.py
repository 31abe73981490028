#!/usr/bin/env python
"""Small batches: one configuration per thread against G lanes per configuration (BRBD_COOP_MAX_BATCH=0 forces the first,
a large value the second), RNEA and ABA, device-resident, CUDA-event time of the kernel (median of 20).
    BRBD_COOP_MAX_BATCH=0 python scripts/small_batch_sweep.py ; BRBD_COOP_MAX_BATCH=1000000 python scripts/small_batch_sweep.py"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import pinocchio_b200 as pb
from conftest import load_model, random_inputs
mode = os.environ.get("BRBD_COOP_MAX_BATCH", "default")
for name in ("talos_reduced_ff", "humanoid_random", "manipulator"):
    model = load_model(name); pool = pb.ModelPool(model, [0])
    pool.set_stream(torch.cuda.current_stream().cuda_stream)
    q, v, a = random_inputs(model, 1 << 15, 1)
    for B in (256, 1024, 2048, 4096, 8192, 16384, 32768):
        tq, tv, ta = (torch.from_numpy(np.ascontiguousarray(x[:, :B].T)).cuda() for x in (q, v, a))
        out = torch.empty((B, model.nv), dtype=torch.float64, device="cuda")
        for algo, fn in (("rnea", lambda: pb.rneaInParallel(1, pool, tq, tv, ta, out, async_=True)),
                         ("aba", lambda: pb.abaInParallel(1, pool, tq, tv, ta, out, async_=True))):
            for _ in range(3): fn()
            torch.cuda.synchronize()
            ts = []
            for _ in range(20):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
            us = float(np.median(ts)) * 1e3
            print(json.dumps({"coop_max_batch": mode, "model": name, "algo": algo, "batch": B, "us": round(us, 1), "configs_per_s": B / (us * 1e-6)}), flush=True)
    pool.close()
