#!/usr/bin/env python
"""Which caller layouts the TMA column store of CRBA accepts: each case in its own process (a faulting launch kills the context)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = {"A_pairs_only": ("humanoid", 1), "B_xoff_only": ("simple_humanoid_ff", 1), "C_both": ("simple_humanoid_ff", 0),
         "D_neither_padded": ("humanoid", 2), "E_neither": ("humanoid", 0)}
if len(sys.argv) > 1:
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np, torch
    import pinocchio_b200 as pb
    from conftest import load_model, random_inputs
    name, pad = CASES[sys.argv[1]]
    model = load_model(name)
    pool = pb.ModelPool(model, [0])
    B, nn = 77, model.nv * model.nv
    q, _, _ = random_inputs(model, B, 5)
    tq = torch.from_numpy(np.ascontiguousarray(q.T)).cuda()
    big = torch.full((B, nn + pad), -7.0, dtype=torch.float64, device="cuda")
    M = big[:, :nn]
    pb.crbaInParallel(1, pool, tq, M)
    torch.cuda.synchronize()
    os.environ["BRBD_CRBA_V"] = "tmem"
    big2 = torch.full((B, nn + pad), -7.0, dtype=torch.float64, device="cuda")
    pb.crbaInParallel(1, pool, tq, big2[:, :nn])
    torch.cuda.synchronize()
    print(sys.argv[1], "ld", nn + pad, "equal:", bool(torch.equal(big, big2)), "maxdiff", float((big - big2).abs().max()))
else:
    for c in CASES:
        r = subprocess.run([sys.executable, __file__, c], capture_output=True, text=True)
        print(r.stdout.strip() or (c + " FAILED: " + r.stderr.strip().splitlines()[-1][:200]))
