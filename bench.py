#!/usr/bin/env python
"""bench.py — headline benchmark of the batched-dynamics hot path (BASELINE.json metric).

A "step" is one pass of the hot path over one batch of synthetic input.  Workload at every N
(weak scaling, per-GPU work fixed): BASELINE.json configs[1] —
    ABA + CRBA on a batch of 65536 configurations of simple_humanoid.urdf + free-flyer, FP64
(one eval = one configuration through one algorithm, so one step = 2 * 65536 evals per GPU).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N ...             # CPU arm: the OpenMP restatement of
                                                              # rneaInParallel/abaInParallel (oracle/)
Prints ONE JSON line (rank 0).  `value` = device-resident throughput (inputs already in HBM, CUDA
events on the launching stream, max over ranks); `e2e` = same metric through the public host-pointer
API with pinned HOST buffers (H2D + kernels + D2H inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

MODEL = "simple_humanoid_ff"
BATCH = 65536
L2_BYTES = 126 * 1024 * 1024
# dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed `ncu --set full` captures
# (profiles/r1_v8_step_ncu_full.csv: crba_tma_kernel<double,224,true> 47.3 MB read + 596.1 MB written; aba_rr_kernel<double,224>
# 388.1 MB read + 426.1 MB written — the ABA figure is 11x its 74 MB of algorithmic bytes: the per-thread pass-3 record
# store, 186 MB for a resident wave, does not stay L2-resident); bytes per launch of 65536 configurations
NCU_TRAFFIC = {"crba": 643.4e6, "aba": 814.2e6}


def load_model(name):
    from pinocchio_b200.model import Model
    with open(os.path.join(ROOT, "tests", "golden", "models", name + ".json")) as fh:
        return Model.from_json(fh.read())


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        self.join(timeout=2)
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def algorithmic_numbers(model, orc):
    """Per-configuration algorithmic bytes (exact) and FLOPs (counted by the oracle's
    counting scalar on one random configuration)."""
    from conftest import random_inputs
    nq, nv = model.nq, model.nv
    q, v, a = random_inputs(model, 1, 3)
    fl = {k: orc.count_flops(k, q[:, 0], v[:, 0], a[:, 0]) for k in ("aba", "crba_world")}
    return {
        "aba": {"bytes": 8 * (nq + 3 * nv), "flops": fl["aba"]["flops"], "sincos": fl["aba"]["sincos"]},
        "crba": {"bytes": 8 * (nq + nv * nv), "flops": fl["crba_world"]["flops"], "sincos": fl["crba_world"]["sincos"]},
    }


def host_threads():
    """Every hardware thread this process may run on.  Not omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1 to
    its workers, which would time the CPU arm on one core."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_sample(orc, model, nthreads, target_s=12.0):
    """ABA + CRBA with the OpenMP oracle on a bounded sample of the workload; returns (evals/s, sample size)."""
    from conftest import random_inputs
    n = 2048
    q, v, tau = random_inputs(model, n, 77)
    t0 = time.perf_counter()
    orc.aba(q, v, tau, nthreads=nthreads)
    orc.crba(q, nthreads=nthreads, world=True)
    probe = time.perf_counter() - t0
    n2 = int(min(BATCH, max(n, n * target_s / max(probe, 1e-6))))
    q, v, tau = random_inputs(model, n2, 78)
    orc.aba(q[:, :256], v[:, :256], tau[:, :256], nthreads=nthreads)  # warm-up
    t0 = time.perf_counter()
    orc.aba(q, v, tau, nthreads=nthreads)
    orc.crba(q, nthreads=nthreads, world=True)
    dt = time.perf_counter() - t0
    return 2.0 * n2 / dt, n2, dt


def run_reference(args):
    """CPU arm: the restated reference path (oracle port) with all host threads, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import Oracle, build_oracle
    build_oracle()
    model = load_model(MODEL)
    orc = Oracle(model)
    nthreads = host_threads()
    from conftest import random_inputs
    n = 4096
    q, v, tau = random_inputs(model, n, 5)
    times = []
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        orc.aba(q, v, tau, nthreads=nthreads)
        orc.crba(q, nthreads=nthreads, world=True)
        dt = time.perf_counter() - t0
        if s >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = 2.0 * n / (ms * 1e-3)
    line = {
        "impl": "reference", "metric": "batched dynamics evals/sec (ABA + CRBA)", "value": value, "unit": "evals/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"ABA + CRBA, {MODEL} (simple_humanoid.urdf + free-flyer, nq=36 nv=35), FP64; "
                               f"CPU arm times a bounded sample of {n} configurations per step"},
        "cpu_baseline": {"value": value, "unit": "evals/s", "cores": nthreads, "kind": "port",
                         "sample": f"{n} configurations per step through ABA(WORLD) + CRBA(WORLD), OpenMP schedule(static), "
                                   f"the restated rneaInParallel/abaInParallel driver (oracle/), {nthreads} threads"},
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="configurations per GPU per step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import pinocchio_b200 as pb
    from conftest import random_inputs

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)

    model = load_model(MODEL)
    nq, nv, B = model.nq, model.nv, args.batch
    pool = pb.ModelPool(model, [local_rank])
    stream = torch.cuda.current_stream()
    pool.set_stream(stream.cuda_stream)

    # Rotating input sets so that consecutive steps never find their inputs in L2 (126 MB):
    in_bytes = 8 * B * (nq + 2 * nv)
    nsets = max(2, int(np.ceil(2.5 * L2_BYTES / in_bytes)))
    sets = []
    for s in range(nsets):
        q, v, tau = random_inputs(model, B, 1000 * rank + 10 * s)
        sets.append(tuple(torch.from_numpy(np.ascontiguousarray(x.T)).to(dev) for x in (q, v, tau)))
    a_out = torch.empty((B, nv), dtype=torch.float64, device=dev)
    M_out = torch.empty((B, nv * nv), dtype=torch.float64, device=dev)

    def step(i, ev=None):
        q, v, tau = sets[i % nsets]
        if ev is not None:
            ev[0].record(stream)
        pb.abaInParallel(1, pool, q, v, tau, a_out, async_=True)
        if ev is not None:
            ev[1].record(stream)
        pb.crbaInParallel(1, pool, q, M_out, async_=True)
        if ev is not None:
            ev[2].record(stream)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    barrier()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    launches0 = pool.launch_count()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for i in range(args.steps):
        step(args.warmup + i, evs[i])
    e1.record(stream)
    barrier()
    launches = pool.launch_count() - launches0
    total_ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if sampler else None
    aba_ms = float(np.mean([e[0].elapsed_time(e[1]) for e in evs]))
    crba_ms = float(np.mean([e[1].elapsed_time(e[2]) for e in evs]))
    t = torch.tensor([total_ms, aba_ms, crba_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, aba_ms, crba_ms = (float(x) for x in t.cpu())
    ms_per_step = total_ms / args.steps
    value = 2.0 * B * world / (ms_per_step * 1e-3)

    # ---- end to end through the public API with pinned HOST buffers (H2D + kernels + D2H timed) ----
    pool.set_stream(None)
    e2e_steps = max(3, min(args.steps, 5))
    hq, hv, ht = (torch.from_numpy(np.ascontiguousarray(x.cpu().numpy())).pin_memory() for x in sets[0])
    ha = torch.empty((B, nv), dtype=torch.float64).pin_memory()
    hM = torch.empty((B, nv * nv), dtype=torch.float64).pin_memory()
    nq_, nv_ = nq, nv
    hq_n, hv_n, ht_n, ha_n, hM_n = (x.numpy().T for x in (hq, hv, ht, ha, hM))  # (rows x B) column-major views

    def e2e_step():
        pb.abaInParallel(1, pool, hq_n, hv_n, ht_n, ha_n)
        pb.crbaInParallel(1, pool, hq_n, hM_n)
        return float(ha_n[0, 0]) + float(hM_n[0, 0])

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    tt = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_s = float(tt.cpu()[0])
    e2e_value = 2.0 * B * world * e2e_steps / e2e_s
    h2d = 8 * B * ((nq + 2 * nv) + nq)   # ABA inputs + CRBA input
    d2h = 8 * B * (nv + nv * nv)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel + CPU baseline (rank 0) ----------------------------------
    from oracle import Oracle, build_oracle
    build_oracle()
    orc = Oracle(model)
    alg = algorithmic_numbers(model, orc)
    hbm_peak, peak_src = measured_peaks()
    fp64_peak, _ = pool.measure_fp64_peak()
    kern = {}
    for name, ms in (("aba", aba_ms), ("crba", crba_ms)):
        gbs = alg[name]["bytes"] * B / (ms * 1e-3) / 1e9
        tfl = alg[name]["flops"] * B / (ms * 1e-3) / 1e12
        kern[name] = {"ms_per_launch": ms, "configs_per_s": B / (ms * 1e-3), "algorithmic_bytes_per_config": alg[name]["bytes"],
                      "algorithmic_flops_per_config": alg[name]["flops"], "sincos_per_config": alg[name]["sincos"],
                      "achieved_GBs": gbs, "hbm_frac": gbs / hbm_peak, "achieved_fp64_TFLOPs": tfl,
                      "fp64_frac_of_measured_dfma_peak": tfl / (fp64_peak / 1e12)}
    # The step has two kernels.  CRBA is HBM-bound (AI 1.5 flop/B) and is the kernel the `roofline` object (whose
    # `bound` is "hbm" | "tensor") describes; ABA is bound by the FP64 pipe (AI 24 flop/B, no tensor cores on this
    # path) and is reported against the measured DFMA peak under `fp64_roofline`.  `share_of_step` says how the
    # step's time splits, so the dominant kernel can be read off either way.
    kname = {"aba": "aba_rr_kernel<double>", "crba": "crba_tma_kernel<double>"}
    roofline = {"bound": "hbm", "kernel": kname["crba"], "achieved": kern["crba"]["achieved_GBs"], "peak": hbm_peak,
                "unit": "GB/s", "frac": kern["crba"]["hbm_frac"], "traffic": NCU_TRAFFIC.get("crba"), "peak_source": peak_src,
                "share_of_step": crba_ms / (aba_ms + crba_ms)}
    fp64_roofline = {"bound": "fp64", "kernel": kname["aba"], "achieved": kern["aba"]["achieved_fp64_TFLOPs"],
                     "peak": fp64_peak / 1e12, "unit": "TFLOP/s", "frac": kern["aba"]["fp64_frac_of_measured_dfma_peak"],
                     "traffic": NCU_TRAFFIC.get("aba"), "peak_source": "measured in this run (register-resident DFMA loop)",
                     "share_of_step": aba_ms / (aba_ms + crba_ms)}
    line = {
        "metric": "batched dynamics evals/sec (ABA + CRBA)", "value": value, "unit": "evals/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"ABA + CRBA on {B} configurations per GPU of {MODEL} (simple_humanoid.urdf + free-flyer, "
                               f"nq={nq} nv={nv}), FP64; one eval = one configuration through one algorithm",
                   "batch_per_gpu": B, "l2_policy": f"inputs rotate over {nsets} resident sets ({nsets * in_bytes >> 20} MiB > 126 MiB L2); "
                                                    f"the {8 * B * nv * nv >> 20} MiB CRBA output streams through L2 every step",
                   "parallelism": f"batch sharded over {world} GPU(s), no collective"},
        "clocks": clocks, "gpu_launches": int(launches),
        "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "steps": e2e_steps, "note": "pinned host buffers -> brbd_*_batch(BRBD_PTR_HOST): H2D + kernels + D2H, wall clock"},
        "roofline": roofline,
        "fp64_roofline": fp64_roofline,
        "kernels": kern,
        "fp64_peak_measured_TFLOPs": fp64_peak / 1e12,
    }
    if world == 1 and not args.no_cpu:
        nthreads = host_threads()
        cpu_v, n2, dt = cpu_sample(orc, model, nthreads)
        line["cpu_baseline"] = {"value": cpu_v, "unit": "evals/s", "cores": nthreads, "kind": "port",
                                "sample": f"{n2} configurations through ABA(WORLD) + CRBA(WORLD) in {dt:.1f} s with the OpenMP "
                                          f"restatement of rneaInParallel/abaInParallel (oracle/), schedule(static)"}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
