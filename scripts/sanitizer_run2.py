#!/usr/bin/env python
"""The kernels added in the second half of round 2, one small call each, for compute-sanitizer: generated CRBA (bulk copies /
compact staging / packed output), computeMinverse by Cholesky (simple loops for a small model, 4 x 4 blocks for a humanoid),
generated derivative kernels with bulk-copy output, the packed transfer with host threads."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import pinocchio_b200 as pb
from conftest import load_model, random_inputs
for name in (sys.argv[1:] or ["simple_humanoid_ff", "manipulator"]):
    model = load_model(name)
    q, v, a = random_inputs(model, 70, 1)
    for mode in ("bulk", "lsu", "compact"):
        os.environ["BRBD_GEN_CRBA_MODE"] = mode
        pool = pb.ModelPool(model, [0]); pool.specialize(["crba"], min_batch=1)
        pb.crbaInParallel(1, pool, q); print(name, "generated crba", mode, flush=True)
        pool.close()
    del os.environ["BRBD_GEN_CRBA_MODE"]
    pool = pb.ModelPool(model, [0])
    pb.crbaPackedInParallel(1, pool, q); print(name, "packed crba", flush=True)
    pb.computeMinverseInParallel(1, pool, q); print(name, "Minv (Cholesky)", flush=True)
    qq, _, _ = random_inputs(model, 4200, 2)
    pb.crbaInParallel(3, pool, qq); print(name, "crba, packed transfer + host threads", flush=True)
    if model.nv <= 16:
        pool.specialize(["rnea_derivatives", "aba_derivatives"], min_batch=1)
        pb.computeRNEADerivativesInParallel(1, pool, q, v, a); pb.computeABADerivativesInParallel(1, pool, q, v, a)
        print(name, "generated derivatives", flush=True)
    pool.close()
print("done")
