#!/usr/bin/env python
"""bench.py -- the headline metric of BASELINE.json on B200:

    log-mll + gradient (update_mll_and_dmll!, src/GPE.jl:332-335) for GPE SEIso, N=32768, d=8, FP64
    reported as algorithmic GFLOP/s (F_alg = N^3 + 2 N^2) and ms per evaluation.

A "step" is one evaluation of the hot path (Gram build -> Cholesky -> alpha/mll -> K_y^-1 ->
fused gradient trace) on synthetic inputs (x, y ~ N(0,1), seed 1; SEIso(0.3,0.3), logNoise 0.3,
MeanConst(0) -- config C2 of SURVEY.md §8(d)).

  value        inputs resident in HBM, timed with CUDA events on the engine's stream
  e2e          same step through the public GPE API with HOST buffers: x, y uploaded from pinned host
               memory and alpha / mll / dmll read back inside the timed region
  roofline     the dominant kernel (FP64 DMMA NT GEMM): algorithmic flops / sum of its launch
               durations (CUDA events around every launch in one extra profiled step) vs the DMMA
               issue rate measured in the same run (MEASURED_PEAKS.json has no FP64 figure)
  cpu_baseline the reference's algorithm (oracle port: scalar loops in C + LAPACK via OpenBLAS, all
               host cores) on a bounded sample (N=8192) of the same workload
  --impl reference   times only that CPU path (Julia is not installed in this image).
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gaussianprocesses.jl_b200"))

N_FULL, D = 32768, 8
LL, LSIG, LNOISE = 0.3, 0.3, 0.3
N_CPU_SAMPLE = int(os.environ.get("GPB200_CPU_SAMPLE_N", "8192"))      # bounded CPU sample (test hook: smaller N)
METRIC = "log-mll+grad GFLOP/s, GPE SEIso N=32768 d=8 FP64 (update_mll_and_dmll!)"


def falg(n):
    return float(n) ** 3 + 2.0 * float(n) ** 2


def synth(n, d, seed=1):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((n, d)), rng.standard_normal(n)


class StdoutToStderr:
    """NCCL prints its version banner on the process's C stdout when the first communicator is created; route fd 1 to
    stderr for that stretch so that rank 0's stdout carries nothing but the one JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        try:
            sys.stdout.flush()
        finally:
            os.dup2(self._saved, 1)
            os.close(self._saved)
        return False


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [t.strip() for t in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s in sm if s > 0.5 * (max(mx) if mx else 1)] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline_run(n, steps=1):
    from oracle import cpu_baseline as cb
    X, y = synth(n, D, seed=1)
    best, last = None, None
    for _ in range(steps):
        r = cb.seiso_mll_and_dmll(X, y, LL, LSIG, LNOISE, 0.0)
        t = r["seconds"]["total"]
        best = t if best is None else min(best, t)
        last = r
    return best, last, cb.host_threads()


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU algorithm (oracle port; Julia absent) on host cores."""
    if rank != 0:
        return
    from oracle import cpu_baseline as cb
    n = N_CPU_SAMPLE
    X, y = synth(n, D, seed=1)
    for _ in range(max(args.warmup, 0)):
        cb.seiso_mll_and_dmll(X, y, LL, LSIG, LNOISE, 0.0)
        break                               # one warm-up is enough for a CPU path (thread pools, page faults)
    ts = []
    for _ in range(args.steps):
        ts.append(cb.seiso_mll_and_dmll(X, y, LL, LSIG, LNOISE, 0.0)["seconds"]["total"])
    t = float(np.mean(ts))
    val = falg(n) / t * 1e-9
    cores = cb.host_threads()
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2 sample: GPE SEIso(0.3,0.3) logNoise 0.3 MeanConst(0), d=8, N=%d of 32768 "
                               "(CPU path needs ~5 min and 34 GB at full N)" % n},
        "cpu_baseline": {"value": val, "unit": "GFLOP/s", "cores": cores, "kind": "port",
                         "sample": "N=%d d=8, full update_mll_and_dmll! (scalar cov!/dmll_kern! loops in C, 1 thread; "
                                   "dpotrf + dpotrs(-I) via OpenBLAS on %d threads); GFLOP/s counts F_alg=N^3+2N^2" % (n, cores)},
        "e2e": {"value": val, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=N_FULL, help="override N (debug only; the metric is N=32768)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    import gpb200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback in libgpb200)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        with StdoutToStderr():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()

    N = args.n
    X, y = synth(N, D, seed=1)
    # pinned host staging for the e2e leg
    xp = torch.empty((N, D), dtype=torch.float64).pin_memory(); xp.numpy()[:] = X
    yp = torch.empty((N,), dtype=torch.float64).pin_memory(); yp.numpy()[:] = y
    Xh, yh = xp.numpy(), yp.numpy()

    gp = gpb200.GPE(Xh.T, yh, gpb200.MeanConst(0.0), gpb200.SEIso(LL, LSIG), LNOISE, device=local_rank)
    eng = gp._eng
    if world > 1:
        # strong scaling: the SAME N=32768 problem, work partitioned over the ranks (block-column
        # Cholesky with NCCL panel broadcasts, split inverse, row-cyclic W'W + trace)
        with StdoutToStderr():
            gp.init_distributed()
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    theta = np.array([LL, LSIG])
    r = yh - 0.0

    def step_resident():
        eng.factorize(theta, LNOISE)
        eng.mll(r)
        eng.grad_prepare()
        return eng.grad_kernel()

    def step_e2e():
        gp.reload_data(Xh.T, yh)            # H2D of x and (inside update_mll) y
        gp.update_mll_and_dmll()            # D2H alpha, mll, dmll
        return gp.mll

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_resident()
    sampler = ClockSampler(local_rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    l0 = eng.launch_count()
    sampler.start()
    e0.record(stream)
    for _ in range(args.steps):
        step_resident()
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    launches = eng.launch_count() - l0
    ms_total = e0.elapsed_time(e1)
    tmr = eng.timings()
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps

    # ---- e2e: host buffers in, host results out, through the public GPE API -------------------
    step_e2e()
    barrier()
    t0 = time.perf_counter()
    e0.record(stream)
    for _ in range(args.steps):
        step_e2e()
    e1.record(stream)
    barrier()
    e2e_wall = (time.perf_counter() - t0) * 1e3 / args.steps
    e2e_ms = max(e0.elapsed_time(e1) / args.steps, e2e_wall)
    if world > 1:
        t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    h2d = X.nbytes + y.nbytes + theta.nbytes + 8
    d2h = 8 * N + 8 + 8 * (theta.size + 1)

    # ---- roofline of the dominant kernel: one profiled step (events around every GEMM launch) ----
    roof = None
    pk = eng.fp64_peak()
    eng.set_option("profile", 1)
    step_resident()
    torch.cuda.synchronize()
    tp = eng.timings()
    eng.set_option("profile", 0)
    if tp["gemm_ms"] > 0:
        alg = float(N) ** 3 / world                             # N^3/3 Cholesky + 2N^3/3 inverse, all in this kernel (per rank)
        achieved = alg / (tp["gemm_ms"] * 1e-3) * 1e-12
        traffic = None
        tj = os.path.join(ROOT, "profiles", "gemm_traffic.json")
        if os.path.exists(tj):
            try:
                traffic = json.load(open(tj)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        roof = {"bound": "tensor", "kernel": "gpb200_dgemm_nt_tma (FP64 DMMA.8x8x4 fed by TMA)",
                "achieved": achieved, "peak": pk["dmma_tflops"], "unit": "TFLOP/s", "frac": achieved / pk["dmma_tflops"],
                "traffic": traffic,
                "peak_source": "DMMA.8x8x4 issue rate measured in this run (gpb200_fp64_peak); MEASURED_PEAKS.json "
                               "carries no FP64 figure -- its HBM/bf16 numbers do not bound this kernel; DFMA rate "
                               "%.1f TFLOP/s, cuBLAS DGEMM on this pool 36.5 TFLOP/s (profiles/r01_cublas_dgemm.txt)" % pk["dfma_tflops"],
                "launches_per_step": int(tp["gemm_launches"]), "kernel_ms_per_step": tp["gemm_ms"],
                "kernel_share_of_step": tp["gemm_ms"] / ms_step,
                "algorithmic_flops_per_step": alg, "executed_flops_per_step": tp["gemm_flops"]}

    # predict_f timing (batched, M = 4096), reported beside the metric
    Xs = np.random.default_rng(2).standard_normal((4096, D))
    eng.predict(Xs)                        # warm-up: sizes the cross-Gram workspace
    t0 = time.perf_counter()
    eng.predict(Xs)
    predict_ms = (time.perf_counter() - t0) * 1e3

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            tcpu, res, cores = cpu_baseline_run(N_CPU_SAMPLE, steps=1)
            cpu = {"value": falg(N_CPU_SAMPLE) / tcpu * 1e-9, "unit": "GFLOP/s", "cores": cores, "kind": "port",
                   "sample": "N=%d d=8 (of 32768): one full update_mll_and_dmll! with the reference's algorithm -- scalar "
                             "cov!/dmll_kern! loops (C, 1 thread) + dpotrf + dpotrs(-I) (OpenBLAS, %d threads); %.1f s; "
                             "phases %s" % (N_CPU_SAMPLE, cores, tcpu, {k: round(v, 2) for k, v in res["seconds"].items()})}
        val = falg(N) / (ms_step * 1e-3) * 1e-9
        line = {
            "metric": METRIC, "value": val, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C2: GPE SEIso(0.3,0.3) logNoise 0.3 MeanConst(0), N=%d d=8 FP64: Gram + Cholesky + "
                                   "alpha/mll + K^-1 + gradient trace per step" % N,
                       "l2": "working set (two 8.6 GB N x N FP64 matrices) exceeds the 126 MB L2; no flush needed",
                       "parallelism": "1 GPU" if world == 1 else
                       "%d GPUs: 1-D block-cyclic block columns (NCCL panel broadcast over NVLink, look-ahead), "
                       "split level-parallel inverse + all-gather, tile-row-cyclic W'W/trace + all-reduce of P+1 sums; "
                       "F/G replicated per GPU" % world,
                       "phases_ms": {k: round(v, 3) for k, v in tmr.items() if k in ("gram", "cholesky", "solve_mll", "inverse", "trace")},
                       "predict_f_ms_M4096": predict_ms},
            "clocks": clocks,
            "e2e": {"value": falg(N) / (e2e_ms * 1e-3) * 1e-9, "unit": "GFLOP/s", "ms_per_step": e2e_ms,
                    "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h)},
            "gpu_launches": int(launches),
            "roofline": roof,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
