#!/usr/bin/env python
"""One small call of every kernel family (B = 64) for compute-sanitizer (memcheck / racecheck): the warp-cooperative kernels
lean on __syncwarp phase logic, the generated ones on tcgen05 / cp.async staging.  Prints one line per call."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import pinocchio_b200 as pb
from conftest import load_model, make_extra_models, random_inputs
extra = make_extra_models()
for name in (sys.argv[1:] or ["manipulator", "simple_humanoid_ff", "mixed"]):
    model = extra[name] if name in extra else load_model(name)
    pool = pb.ModelPool(model, [0])
    q, v, a = random_inputs(model, 64, 1)
    for coop in ("1000000", "0"):
        os.environ["BRBD_COOP_MAX_BATCH"] = coop
        pb.rneaInParallel(1, pool, q, v, a); pb.abaInParallel(1, pool, q, v, a)
        print(name, "rnea / aba", "coop" if coop != "0" else "thread", flush=True)
    pb.crbaInParallel(1, pool, q); print(name, "crba", flush=True)
    pb.computeRNEADerivativesInParallel(1, pool, q, v, a); print(name, "rnea derivatives", flush=True)
    pb.computeABADerivativesInParallel(1, pool, q, v, a); print(name, "aba derivatives", flush=True)
    pb.computeMinverseInParallel(1, pool, q); pb.integrateInParallel(1, pool, q, v); pb.abaEulerStepInParallel(1, pool, q, v, a, 1e-3)
    print(name, "Minv / integrate / euler", flush=True)
    pool.specialize(["rnea", "aba", "crba"], min_batch=1)
    os.environ["BRBD_CRBA_V"] = "gen"
    pb.rneaInParallel(1, pool, q, v, a); pb.abaInParallel(1, pool, q, v, a); pb.crbaInParallel(1, pool, q)
    del os.environ["BRBD_CRBA_V"]
    print(name, "generated rnea / aba / crba", flush=True)
    pool.close()
print("done")
