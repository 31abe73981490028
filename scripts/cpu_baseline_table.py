#!/usr/bin/env python
"""The reference's own CPU-runnable case (BASELINE.json configs[0], benchmark/timings-parallel.cpp:141-207) with the restated
CPU path (oracle/, OpenMP schedule(static) as algorithm/parallel/rnea.hpp:74-82), next to the GPU path on the same batches:
  * single-thread microseconds per call for RNEA / ABA / CRBA on talos (the README chart: 4.2 / 8.5 / 5 us on a 2.4 GHz i7);
  * rneaInParallel / abaInParallel at B = 256 (the reference's setting) and B = 1024 (configs[0]) over 1, 2, 4, ... host threads;
  * the same calls through the C ABI from host memory (end to end) and device-resident, when a GPU is present.
One JSON line per measurement.  SURVEY.md §8d: label = "restated CPU baseline", not "Pinocchio"."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_model, random_inputs
from oracle import Oracle, build_oracle
build_oracle()
ncpu = len(os.sched_getaffinity(0))
cpu_model = next((l.split(":", 1)[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")), "?")


def timed(fn, min_s=1.0):
    fn()
    n, t0 = 0, time.perf_counter()
    while True:
        fn(); n += 1
        dt = time.perf_counter() - t0
        if dt >= min_s:
            return dt / n


for name in ("talos_reduced_ff", "humanoid_random"):
    m = load_model(name); o = Oracle(m)
    q, v, a = random_inputs(m, 1024, 1)
    for algo, fn in (("rnea", lambda n, t: o.rnea(q[:, :n], v[:, :n], a[:, :n], nthreads=t)),
                     ("aba", lambda n, t: o.aba(q[:, :n], v[:, :n], a[:, :n], nthreads=t)),
                     ("crba", lambda n, t: o.crba(q[:, :n], nthreads=t, world=True))):
        s = timed(lambda: fn(256, 1))
        print(json.dumps({"what": "single-thread us per call (restated CPU baseline)", "model": name, "algo": algo, "us_per_call": s / 256 * 1e6,
                          "cpu": cpu_model}), flush=True)
        if algo == "crba":
            continue
        for B in (256, 1024):
            t = 1
            while t <= ncpu:
                s = timed(lambda: fn(B, t), 0.5)
                print(json.dumps({"what": "restated rneaInParallel/abaInParallel (OpenMP)", "model": name, "algo": algo, "batch": B, "threads": t,
                                  "configs_per_s": B / s, "us_per_batch": s * 1e6}), flush=True)
                t *= 2
try:
    import torch
    have_gpu = torch.cuda.is_available()
except Exception:
    have_gpu = False
if have_gpu:
    import pinocchio_b200 as pb
    for name in ("talos_reduced_ff", "humanoid_random"):
        m = load_model(name); pool = pb.ModelPool(m, [0])
        q, v, a = random_inputs(m, 1024, 1)
        for B in (256, 1024):
            qh, vh, ah = (pb.pin_host(np.asfortranarray(x[:, :B])) for x in (q, v, a))
            out = pb.pin_host(np.zeros((m.nv, B), order="F"))
            tq, tv, ta = (torch.from_numpy(np.ascontiguousarray(x[:, :B].T)).cuda() for x in (q, v, a))
            tout = torch.empty((B, m.nv), dtype=torch.float64, device="cuda")
            for algo, host_fn, dev_fn in (("rnea", lambda: pb.rneaInParallel(1, pool, qh, vh, ah, out), lambda: pb.rneaInParallel(1, pool, tq, tv, ta, tout)),
                                          ("aba", lambda: pb.abaInParallel(1, pool, qh, vh, ah, out), lambda: pb.abaInParallel(1, pool, tq, tv, ta, tout))):
                s = timed(host_fn, 0.5)
                print(json.dumps({"what": "B200 through the C ABI, pinned host buffers (end to end)", "model": name, "algo": algo, "batch": B,
                                  "configs_per_s": B / s, "us_per_batch": s * 1e6}), flush=True)
                s = timed(dev_fn, 0.5)
                print(json.dumps({"what": "B200 through the C ABI, device-resident (synchronous call)", "model": name, "algo": algo, "batch": B,
                                  "configs_per_s": B / s, "us_per_batch": s * 1e6}), flush=True)
            for x in (qh, vh, ah, out):
                pb.unpin_host(x)
        pool.close()
