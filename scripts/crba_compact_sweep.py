#!/usr/bin/env python
"""Generated CRBA with compact staging: sweep of (entries per flush / 8, threads per CTA) in one process, dense and packed
output, parity against the oracle on the first 1024 configurations, device-resident timing.
    python scripts/crba_compact_sweep.py [--batch 65536] [--configs "4:640,5:640"] [models...]"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import pinocchio_b200 as pb
from conftest import load_model, make_extra_models, random_inputs
from oracle import Oracle
ap = argparse.ArgumentParser()
ap.add_argument("models", nargs="*", default=["simple_humanoid_ff", "talos_reduced_ff"])
ap.add_argument("--batch", type=int, default=65536)
ap.add_argument("--reps", type=int, default=20)
ap.add_argument("--configs", default="3:512,3:768,3:1024,4:512,4:640,4:768,4:896,5:512,5:640,5:704,6:512,6:608,8:384,8:480")
ap.add_argument("--packed", action="store_true")  # also time brbd_crba_packed_batch (its own kernel) beside every config
ap.add_argument("--baseline", action="store_true")  # also time the hand-written kernel and the non-compact generated one
args = ap.parse_args()
extra = make_extra_models()


def timed(fn, reps):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name in args.models:
    model = extra[name] if name in extra else load_model(name)
    orc = Oracle(model)
    B, nv = args.batch, model.nv
    q, _, _ = random_inputs(model, B, 1)
    tq = torch.from_numpy(np.ascontiguousarray(q.T)).cuda()
    ref = orc.crba(q[:, :1024], world=True)
    scale = np.abs(ref).max()
    Md = torch.empty((B, nv * nv), dtype=torch.float64, device="cuda")
    if args.baseline:
        os.environ.pop("BRBD_CRBA_V", None)
        pool = pb.ModelPool(model, [0]); pool.set_stream(torch.cuda.current_stream().cuda_stream)
        ms = timed(lambda: pb.crbaInParallel(1, pool, tq, Md, async_=True), args.reps)
        print(f"{name} B={B} hand-written (crba_tma_kernel): {ms:.4f} ms = {B*(model.nq+nv*nv)*8/ms/1e6:.0f} GB/s", flush=True)
        pool.close()
    # kind:K:NT[:NBUF], kind in bulk | lsu | compact (a bare K:NT means compact)
    todo = []
    for c in args.configs.split(","):
        f = c.split(":")
        if not c: continue
        if f[0].isdigit(): f = ["compact"] + f
        todo.append((f[0], int(f[1]), int(f[2]), int(f[3]) if len(f) > 3 else 1))
    if args.baseline: todo = [("lsu", 3, 256, 1), ("lsu", 1, 512, 1)] + todo
    for kind, K, NT, NBUF in todo:
        os.environ["BRBD_CRBA_V"] = "gen"
        os.environ["BRBD_GEN_CRBA_MODE"] = kind
        os.environ["BRBD_GEN_CRBA_K"] = str(K); os.environ["BRBD_GEN_CRBA_NT"] = str(NT); os.environ["BRBD_GEN_CRBA_NBUF"] = str(NBUF)
        pool = pb.ModelPool(model, [0]); pool.set_stream(torch.cuda.current_stream().cuda_stream)
        try:
            t0 = time.time(); pool.specialize(["crba"]); ts = time.time() - t0
        except Exception as e:
            print(f"{name} {kind} K={K} NT={NT} NBUF={NBUF}: specialize failed: {str(e)[:200]}", flush=True); pool.close(); continue
        Md.fill_(float("nan"))
        pb.crbaInParallel(1, pool, tq, Md, async_=True); torch.cuda.synchronize()
        got = Md[:1024].cpu().numpy().T
        err = np.abs(got - ref).max() / scale
        ok_all = bool(torch.isfinite(Md).all().item())
        ms = timed(lambda: pb.crbaInParallel(1, pool, tq, Md, async_=True), args.reps)
        line = f"{name} B={B} {kind} K={K} NT={NT} NBUF={NBUF}: dense {ms:.4f} ms = {B*(model.nq+nv*nv)*8/ms/1e6:.0f} GB/s err {err:.1e} finite {ok_all}"
        if kind == "compact" or args.packed:
            rows, cols = pool.crbaPattern()
            nnz = len(rows)
            Pd = torch.full((B, nnz), float("nan"), dtype=torch.float64, device="cuda")
            pb.crbaPackedInParallel(1, pool, tq, Pd, async_=True); torch.cuda.synchronize()
            gp = Pd[:1024].cpu().numpy().T
            errp = np.abs(gp - ref[cols.astype(np.int64) * nv + rows]).max() / scale
            okp = bool(torch.isfinite(Pd).all().item())
            msp = timed(lambda: pb.crbaPackedInParallel(1, pool, tq, Pd, async_=True), args.reps)
            line += f" | packed ({nnz} entries) {msp:.4f} ms = {B*(model.nq+nnz)*8/msp/1e6:.0f} GB/s err {errp:.1e} finite {okp}"
        print(line + f" (specialize {ts:.1f}s)", flush=True)
        pool.close()
