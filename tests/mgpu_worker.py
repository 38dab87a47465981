"""Worker for the multi-GPU parity test: torchrun --nproc-per-node R tests/mgpu_worker.py
Each rank builds the same GPE, joins the NCCL communicator, evaluates mll + gradient + predict_f and
rank 0 compares with the CPU oracle (same tolerances as the single-GPU parity tests)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "gaussianprocesses.jl_b200"))


def main():
    import torch
    import torch.distributed as dist
    import gpb200
    from oracle import gp_oracle as orc

    rank = int(os.environ["RANK"]); local = int(os.environ["LOCAL_RANK"]); world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ok = True
    for (N, d, dist_nb, kern, spec) in [
        (1500, 3, 256, gpb200.SEIso(0.3, 0.1), None),
        (2300, 2, 128, gpb200.Mat32Iso(0.2, 0.1) + gpb200.RQIso(0.4, -0.3, 0.2), None),
        (4000, 4, 1024, gpb200.SEIso(0.4, 0.2), None),
    ]:
        rng = np.random.default_rng(N)
        X = rng.standard_normal((N, d)); y = np.sin(X.sum(axis=1)) + 0.1 * rng.standard_normal(N)
        Xs = rng.standard_normal((50, d))
        gp = gpb200.GPE(X.T, y, gpb200.MeanConst(0.2), kern, -0.5, device=local)
        gp._eng.set_option("dist_nb", dist_nb)
        gp.init_distributed(p2p=True)               # maps peer memory (fused panel broadcast): both panel paths are compared below
        gp.update_target_and_dtarget()
        t_p2p = (gp.mll, gp.dmll.copy())
        gp._eng.set_option("p2p", 0)                # same problem through the NCCL panel broadcast
        gp.update_target_and_dtarget()
        same = abs(gp.mll - t_p2p[0]) <= 1e-12 * abs(gp.mll) and np.allclose(gp.dmll, t_p2p[1], rtol=1e-10, atol=1e-12)
        gp._eng.set_option("p2p", 1)
        gp.update_target_and_dtarget()
        if not same:
            print("MGPU p2p and nccl paths disagree", gp.mll, t_p2p[0], flush=True)
            ok = False
        mu, s2 = gp.predict_f(Xs.T)
        # the same problem on ROW-SHARDED storage (each rank maps only its own block rows; NCCL collectives)
        ref = (gp.mll, gp.dmll.copy(), gp.alpha.copy(), mu.copy(), s2.copy())
        gp._eng.set_option("shard", 1)
        gp.update_target_and_dtarget()
        mu_s, s2_s = gp.predict_f(Xs.T)
        info = gp._eng.storage_info()
        sh_ok = (info["sharded"] and abs(gp.mll - ref[0]) <= 1e-11 * abs(ref[0]) and np.allclose(gp.dmll, ref[1], rtol=1e-8, atol=1e-10)
                 and np.max(np.abs(gp.alpha - ref[2])) <= 1e-10 * np.max(np.abs(ref[2]))
                 and np.max(np.abs(mu_s - ref[3])) <= 1e-10 * np.max(np.abs(ref[3])) and np.max(np.abs(s2_s - ref[4])) <= 1e-10 * np.max(np.abs(ref[4])))
        if not sh_ok:
            print("MGPU rank %d: sharded storage disagrees with replicated: mll %.15g vs %.15g, info %s" % (rank, gp.mll, ref[0], info), flush=True)
            ok = False
        elif rank == 0:
            print("MGPU world=%d N=%d: sharded (rb=%d, %.1f MB of F per rank) == replicated" % (world, N, info["rb"], info["bytes_F"] / 1e6), flush=True)
        gp._eng.set_option("shard", 0)
        gp.update_target_and_dtarget()
        if rank == 0:
            o = orc.mll_and_dmll(kern.spec(), X, y, -0.5, ("MeanConst", 0.2))
            mo, vo = orc.predict_f(kern.spec(), X, o, Xs, ("MeanConst", 0.2))
            e_mll = abs(gp.mll - o["mll"]) / abs(o["mll"])
            e_al = np.max(np.abs(gp.alpha - o["alpha"])) / np.max(np.abs(o["alpha"]))
            e_g = np.max(np.abs(gp.dmll - o["dmll"]) / (np.abs(o["dmll"]) + 1e-10))
            e_mu = np.max(np.abs(mu - mo)) / np.max(np.abs(mo))
            e_v = np.max(np.abs(s2 - vo)) / np.max(np.abs(vo))
            good = e_mll < 1e-10 and e_al < 1e-10 and e_g < 1e-8 and e_mu < 1e-10 and e_v < 1e-10
            print("MGPU world=%d N=%d nb=%d: mll %.2e alpha %.2e dmll %.2e mu %.2e var %.2e -> %s"
                  % (world, N, dist_nb, e_mll, e_al, e_g, e_mu, e_v, "OK" if good else "FAIL"), flush=True)
            ok = ok and good
        # every rank must hold the same results
        t = torch.tensor([gp.mll] + list(gp.dmll), dtype=torch.float64, device="cuda")
        tmax = t.clone(); tmin = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX); dist.all_reduce(tmin, op=dist.ReduceOp.MIN)
        if rank == 0 and not torch.equal(tmax, tmin):
            print("MGPU ranks disagree", (tmax - tmin).abs().max().item(), flush=True)
            ok = False
    # ---- FITC with the observations sharded over the ranks (one all-reduce of the M x M accumulators) against the
    #      single-GPU engine on the full data set ----
    rng = np.random.default_rng(77)
    N, M, d = 6000, 140, 3
    X = rng.standard_normal((N, d)); y = np.sin(X.sum(1)) + 0.1 * rng.standard_normal(N)
    Xu = X[rng.permutation(N)[:M]]
    Xs = rng.standard_normal((60, d))
    kf = gpb200.SEArd([-0.4, -0.6, -0.5], 0.1)
    lo, hi = rank * N // world, (rank + 1) * N // world
    gd = gpb200.FITC(X[lo:hi].T, Xu.T, y[lo:hi], gpb200.MeanConst(0.1), kf, -1.0, device=local, distributed=True)
    gd.update_mll_and_dmll()
    mu_d, s2_d = gd.predict_f(Xs.T)
    gs = gpb200.FITC(X.T, Xu.T, y, gpb200.MeanConst(0.1), kf, -1.0, device=local)          # reference: everything on one GPU
    gs.update_mll_and_dmll()
    mu_s, s2_s = gs.predict_f(Xs.T)
    f_ok = (abs(gd.mll - gs.mll) <= 1e-10 * abs(gs.mll) and np.max(np.abs(gd.alpha - gs.alpha[lo:hi])) <= 1e-9 * np.max(np.abs(gs.alpha))
            and np.allclose(gd.dmll[[0, 2, 3, 4, 5]], gs.dmll[[0, 2, 3, 4, 5]], rtol=1e-7, atol=1e-8)
            and np.max(np.abs(mu_d - mu_s)) <= 1e-9 * np.max(np.abs(mu_s)) and np.max(np.abs(s2_d - s2_s)) <= 1e-9 * np.max(np.abs(s2_s)))
    # the mean-parameter gradient is a sum over the rank's own rows: add the slices up
    t = torch.tensor([gd.dmll[1]], dtype=torch.float64, device="cuda"); dist.all_reduce(t)
    f_ok = f_ok and abs(t.item() - gs.dmll[1]) <= 1e-7 * abs(gs.dmll[1]) + 1e-8
    print("MGPU rank %d FITC sharded over %d ranks: mll %.12f vs %.12f -> %s" % (rank, world, gd.mll, gs.mll, "OK" if f_ok else "FAIL"), flush=True)
    tf = torch.tensor([1.0 if f_ok else 0.0], dtype=torch.float64, device="cuda"); dist.all_reduce(tf, op=dist.ReduceOp.MIN)
    ok = ok and tf.item() > 0.5
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_RESULT", "PASS" if ok else "FAIL", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
