#!/usr/bin/env python
"""Generated computeRNEADerivatives / computeABADerivatives (small models): threads per CTA sweep, device-resident timing.
    python scripts/derivs_gen_sweep.py [--batch 1048576] [--nts 64,96,128,192] [models...]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import pinocchio_b200 as pb
from conftest import load_model, make_extra_models, random_inputs
from oracle import Oracle
ap = argparse.ArgumentParser()
ap.add_argument("models", nargs="*", default=["manipulator"])
ap.add_argument("--batch", type=int, default=1 << 20)
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--nts", default="64,96,128,160,192,224")
args = ap.parse_args()
extra = make_extra_models()
for name in args.models:
    model = extra[name] if name in extra else load_model(name)
    orc = Oracle(model)
    B, nv = args.batch, model.nv
    q, v, x = random_inputs(model, B, 1)
    tq, tv, tx = (torch.from_numpy(np.ascontiguousarray(t.T)).cuda() for t in (q, v, x))
    outs = [torch.empty((B, nv * nv), dtype=torch.float64, device="cuda") for _ in range(3)] + [torch.empty((B, nv), dtype=torch.float64, device="cuda")]
    for nt in [int(t) for t in args.nts.split(",")]:
        os.environ["BRBD_GEN_NT"] = str(nt)
        pool = pb.ModelPool(model, [0]); pool.set_stream(torch.cuda.current_stream().cuda_stream)
        pool.specialize(["rnea_derivatives", "aba_derivatives"])
        line = f"{name} B={B} NT={nt}:"
        for algo, fn, ref_fn in (("rnea_derivatives", pb.computeRNEADerivativesInParallel, orc.rnea_derivatives), ("aba_derivatives", pb.computeABADerivativesInParallel, orc.aba_derivatives)):
            for o in outs: o.fill_(float("nan"))
            fn(1, pool, tq, tv, tx, *outs, async_=True); torch.cuda.synchronize()
            ref = ref_fn(q[:, :512], v[:, :512], x[:, :512])
            err = max(np.abs(o[:512].cpu().numpy().T - r).max() / max(1.0, np.abs(r).max()) for o, r in zip(outs, ref))
            fin = all(bool(torch.isfinite(o).all().item()) for o in outs)
            for _ in range(2): fn(1, pool, tq, tv, tx, *outs, async_=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps): fn(1, pool, tq, tv, tx, *outs, async_=True)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            line += f" {algo} {ms:.4f} ms ({B*(model.nq+3*nv+3*nv*nv)*8/ms/1e6:.0f} GB/s, err {err:.1e}, finite {fin}) |"
        print(line, flush=True)
        pool.close()
