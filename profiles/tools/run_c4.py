"""Config C4 of BASELINE.json / SURVEY.md §8(d): GPE Mat32Iso(log 2, 0), N=131072, d=16, FP64, logNoise 0, MeanZero,
row-sharded Gram + block Cholesky over the ranks (one process per GPU, NCCL over NVLink).  A replicated N x N matrix is
137 GB: the run is possible only because every rank maps its own block rows (csrc/shard_impl.cuh).

  torchrun --nproc-per-node 8 profiles/tools/run_c4.py [--npts 131072] [--dim 16] [--mtest 4096] [--evals 1] [--storage 1]
  (option names avoid torchrun's own prefixes: its argparse would otherwise claim e.g. --n)

Prints one bench-style JSON line (rank 0): mll+grad time (device events, max over ranks), phases, storage per rank, and
size-independent parity properties (the oracle cannot run at this size): residual of K_y alpha = y on sampled rows (rows of
K_y rebuilt on the host from x), directional finite difference of mll against dmll, predictive mean at training points."""
import argparse, json, math, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gaussianprocesses.jl_b200"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--npts", type=int, default=131072)
    ap.add_argument("--dim", type=int, default=16)
    ap.add_argument("--mtest", type=int, default=4096)
    ap.add_argument("--evals", type=int, default=1)
    ap.add_argument("--storage", type=int, default=1)
    ap.add_argument("--rowblock", type=int, default=0)
    ap.add_argument("--fdcheck", type=int, default=1)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import gpb200
    from gpb200.dist import init_engine_comm
    rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    N, d = args.npts, args.dim
    rng = np.random.default_rng(4)
    X = rng.standard_normal((N, d)); y = rng.standard_normal(N)
    Xs = np.random.default_rng(44).standard_normal((args.mtest, d))
    kern = gpb200.Mat32Iso(math.log(2.0), 0.0)
    eng = gpb200.Engine(local)
    if world > 1:
        init_engine_comm(eng, p2p=False)                 # communicator BEFORE the data: storage mode depends on it
    eng.set_option("shard", args.storage)
    if args.rowblock:
        eng.set_option("shard_rb", args.rowblock)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    t0 = time.time()
    gp = gpb200.GPE(X.T, y, gpb200.MeanZero(), kern, 0.0, engine=eng)      # uploads x, first update_mll!
    t_first = time.time() - t0
    info = eng.storage_info()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(v):
        if world == 1:
            return float(v)
        t = torch.tensor([float(v)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    ms_mll, ms_grad, phases = [], [], None
    for _ in range(args.evals):
        barrier()
        e0.record(stream)
        gp.update_mll()
        e1.record(stream)
        gp.update_dmll()
        e2.record(stream)
        barrier()
        ms_mll.append(maxr(e0.elapsed_time(e1))); ms_grad.append(maxr(e1.elapsed_time(e2)))
        phases = eng.timings()
    barrier()
    t0 = time.time()
    mu, s2 = gp.predict_f(Xs.T)
    barrier()
    t_pred = maxr((time.time() - t0) * 1e3)
    t_pred_dev = maxr(eng.timings()["predict"])
    # ---- parity properties ----
    from oracle import gp_oracle as orc
    spec = kern.spec()
    rows = np.random.default_rng(7).choice(N, 48, replace=False)
    Kr = orc.cov(spec, X[rows], X)
    Kr[np.arange(rows.size), rows] += 1.0                   # exp(2 * logNoise), logNoise = 0
    resid = float(np.max(np.abs(Kr @ gp.alpha - y[rows])) / np.max(np.abs(y)))
    mu_tr, _ = gp.predict_f(X[rows].T)
    Kr[np.arange(rows.size), rows] -= 1.0
    pred_err = float(np.max(np.abs(mu_tr - Kr @ gp.alpha)) / np.max(np.abs(mu_tr)))
    var_ok = bool(np.all(s2 >= 0) and np.all(s2 <= 1.0 + 1e-12))
    fd_rel = None
    g0 = gp.dmll.copy(); mll0 = gp.mll
    if args.fdcheck:
        p0 = gp.get_params(); dirv = np.array([0.4, -0.3, 0.5]); h = 1e-4
        gp.set_params(p0 + h * dirv); gp.update_mll(); tp = gp.mll
        gp.set_params(p0 - h * dirv); gp.update_mll(); tm = gp.mll
        gp.set_params(p0); gp.update_mll()
        fd = (tp - tm) / (2 * h)
        fd_rel = float(abs(fd - g0 @ dirv) / abs(fd))
    if rank == 0:
        falg = float(N) ** 3 + 2.0 * float(N) ** 2
        ms = float(np.median(ms_mll) + np.median(ms_grad))
        line = {
            "metric": "log-mll+grad GFLOP/s, GPE Mat32Iso N=%d d=%d FP64 (update_mll_and_dmll!), row-sharded storage" % (N, d),
            "value": falg / (ms * 1e-3) * 1e-9, "unit": "GFLOP/s", "n_gpus": world, "steps": args.evals, "ms_per_step": ms,
            "ms_mll": float(np.median(ms_mll)), "ms_grad": float(np.median(ms_grad)), "higher_is_better": True, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": "C4: GPE Mat32Iso(log 2, 0) logNoise 0 MeanZero, N=%d d=%d: Gram + Cholesky + alpha/mll + K^-1 + trace" % (N, d),
                       "parallelism": "%d ranks, F/G row-sharded block-cyclic (rb=%d tiles), NCCL" % (world, info["rb"]),
                       "phases_ms": {k: round(v, 2) for k, v in phases.items() if k in ("gram", "cholesky", "solve_mll", "inverse", "trace")},
                       "first_update_mll_incl_upload_s": t_first, "predict_f_M%d_ms" % args.mtest: t_pred, "predict_f_dev_ms": t_pred_dev},
            "storage": {"sharded": info["sharded"], "GB_F_per_rank": info["bytes_F"] / 1e9, "GB_G_per_rank": info["bytes_G"] / 1e9,
                        "GB_one_full_matrix": 8.0 * N * N / 1e9, "tma": info["tma"]},
            "check": {"mll": float(mll0), "dmll": [float(v) for v in g0], "alpha_l1": float(np.sum(np.abs(gp.alpha))),
                      "residual_Kalpha_minus_y_rel": resid, "predict_mean_at_training_rows_rel": pred_err, "variance_in_range": var_ok,
                      "fd_directional_rel": fd_rel},
        }
        print(json.dumps(line), flush=True)
        ok = resid <= 1e-10 and pred_err <= 1e-10 and var_ok and (fd_rel is None or fd_rel <= 1e-6)
        print("C4_RESULT", "PASS" if ok else "FAIL", flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
