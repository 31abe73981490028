#!/usr/bin/env python
"""Specialised (generated) vs generic kernels: parity against the oracle and device-resident timing.
    python scripts/gen_quick.py [models...] [--batch 65536]"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import pinocchio_b200 as pb
from conftest import load_model, make_extra_models, random_inputs
from oracle import Oracle
ap = argparse.ArgumentParser()
ap.add_argument("models", nargs="*", default=["simple_humanoid_ff", "talos_reduced_ff", "manipulator"])
ap.add_argument("--batch", type=int, default=65536)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--algos", default="rnea,aba,crba")          # also: crba_packed (brbd_crba_packed_batch)
ap.add_argument("--skip-generic", action="store_true")
args = ap.parse_args()
extra = make_extra_models()
for name in args.models:
    model = extra[name] if name in extra else load_model(name)
    orc = Oracle(model)
    B = args.batch
    q, v, x = random_inputs(model, B, 1)
    tq, tv, tx = (torch.from_numpy(np.ascontiguousarray(t.T)).cuda() for t in (q, v, x))
    res = {}
    algos = args.algos.split(",")
    for mode in (("generated",) if args.skip_generic else ("generic", "generated")):
        pool = pb.ModelPool(model, [0])
        if mode == "generated":
            t0 = time.time(); pool.specialize([a for a in ("rnea", "aba", "crba") if a in algos or (a == "crba" and "crba_packed" in algos)]); print(f"{name}: specialize {time.time()-t0:.1f}s", flush=True)
        pool.set_stream(torch.cuda.current_stream().cuda_stream)
        crba_fn = lambda n, pl, a, b, c, out=None, async_=False: pb.crbaInParallel(n, pl, a, out, async_=async_)
        packed_fn = lambda n, pl, a, b, c, out=None, async_=False: pb.crbaPackedInParallel(n, pl, a, out, async_=async_)
        for algo, fn in (("rnea", pb.rneaInParallel), ("aba", pb.abaInParallel), ("crba", crba_fn), ("crba_packed", packed_fn)):
            if algo not in algos or (algo == "crba_packed" and mode == "generic"): continue
            out = fn(1, pool, tq, tv, tx)
            torch.cuda.synchronize()
            ref = (orc.rnea(q[:, :4096], v[:, :4096], x[:, :4096]) if algo == "rnea" else
                   orc.aba(q[:, :4096], v[:, :4096], x[:, :4096]) if algo == "aba" else orc.crba(q[:, :4096], world=True))
            if algo == "crba_packed":
                rows, cols = pool.crbaPattern()
                ref = ref[cols.astype(np.int64) * model.nv + rows]
            got = out[:4096].cpu().numpy().T
            err = np.abs(got - ref).max() / np.abs(ref).max()
            last = out[-1].cpu().numpy()
            for _ in range(3): fn(1, pool, tq, tv, tx, out, async_=True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.reps): fn(1, pool, tq, tv, tx, out, async_=True)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.reps
            res[(mode, algo)] = ms
            print(f"{name} {mode:9s} {algo:5s} B={B}: {ms:.4f} ms  ({B/ms/1e3:.1f} M cfg/s)  err {err:.1e}", flush=True)
        pool.close()
