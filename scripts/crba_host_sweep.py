#!/usr/bin/env python
"""Host-pointer crbaInParallel (pinned blocks): plain dense copy against the packed transfer + host-thread rebuild, over the
number of host threads and the chunk size (BRBD_EXPAND_CHUNK).   python scripts/crba_host_sweep.py"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import pinocchio_b200 as pb
from conftest import load_model, random_inputs
model = load_model("simple_humanoid_ff")
B = 65536
q, _, _ = random_inputs(model, B, 1)
hq = torch.from_numpy(np.ascontiguousarray(q.T)).pin_memory()
hM = torch.empty((B, model.nv ** 2), dtype=torch.float64).pin_memory()
pool = pb.ModelPool(model, [0]); pool.specialize(["crba"])
def run(threads):
    pb.crbaInParallel(threads, pool, hq.numpy().T, hM.numpy().T)
    t0 = time.perf_counter()
    for _ in range(5): pb.crbaInParallel(threads, pool, hq.numpy().T, hM.numpy().T)
    return (time.perf_counter() - t0) / 5 * 1e3
print(f"dense copy (num_threads = 1): {run(1):.2f} ms", flush=True)
for chunk in (4096, 8192, 16384, 32768):
    os.environ["BRBD_EXPAND_CHUNK"] = str(chunk)
    print(f"chunk {chunk}: " + " | ".join(f"{t} threads {run(t):.2f} ms" for t in (4, 8, 12, 14, 15, 16)), flush=True)
