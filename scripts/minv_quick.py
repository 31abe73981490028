#!/usr/bin/env python
"""computeMinverse: the articulated-body cooperative kernel (BRBD_MINV_V=coop) against crba + Cholesky (default), with the
generic and the generated CRBA underneath; device-resident timing.   python scripts/minv_quick.py [models...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import pinocchio_b200 as pb
from conftest import load_model, random_inputs
models = sys.argv[1:] or ["simple_humanoid_ff", "talos_reduced_ff", "manipulator"]
for name in models:
    model = load_model(name)
    for B in (65536, 1 << 18):
        q, _, _ = random_inputs(model, B, 1)
        tq = torch.from_numpy(np.ascontiguousarray(q.T)).cuda()
        out = torch.empty((B, model.nv ** 2), dtype=torch.float64, device="cuda")
        for mode in ("coop", "chol", "chol+gen"):
            if mode == "coop": os.environ["BRBD_MINV_V"] = "coop"
            else: os.environ["BRBD_MINV_V"] = "chol"
            pool = pb.ModelPool(model, [0]); pool.set_stream(torch.cuda.current_stream().cuda_stream)
            if mode == "chol+gen": pool.specialize(["crba"])
            for _ in range(2): pb.computeMinverseInParallel(1, pool, tq, out, async_=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5): pb.computeMinverseInParallel(1, pool, tq, out, async_=True)
            e1.record(); torch.cuda.synchronize()
            print(f"{name} B={B} computeMinverse[{mode}]: {e0.elapsed_time(e1)/5:.4f} ms", flush=True)
            pool.close()
