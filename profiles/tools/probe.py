"""Quick phase-timing probe (not the bench): python profiles/tools/probe.py N [nb] [gemm]"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "gaussianprocesses.jl_b200"))
import gpb200

N = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 0
gemm = int(sys.argv[3]) if len(sys.argv) > 3 else 0
d = 8
rng = np.random.default_rng(1)
X = rng.standard_normal((N, d)); y = rng.standard_normal(N)
eng = gpb200.Engine(0)
eng.set_data(X)
ops, dims, theta, _ = gpb200.flatten(gpb200.SEIso(0.3, 0.3), d)
eng.set_kernel(ops, dims, theta.size)
eng.set_option("nb", nb); eng.set_option("gemm", gemm)
for rep in range(3):
    t0 = time.time()
    eng.factorize(theta, 0.3)
    alpha, mll = eng.mll(y)
    eng.grad_prepare()
    gk, trA = eng.grad_kernel()
    t1 = time.time()
    tm = eng.timings()
    fl = N ** 3 + 2.0 * N * N
    print("N=%d nb=%d gemm=%d rep=%d wall=%.1f ms  %.2f TFLOP/s  mll=%.6f  " % (N, nb, gemm, rep, (t1 - t0) * 1e3, fl / (t1 - t0) * 1e-12, mll),
          {k: round(v, 2) for k, v in tm.items()}, "launches", eng.launch_count(), flush=True)
if len(sys.argv) > 4:
    M = int(sys.argv[4])
    Xs = rng.standard_normal((M, d))
    for rep in range(2):
        t0 = time.time(); mu, var, _ = eng.predict(Xs); t1 = time.time()
        print("predict M=%d wall=%.1f ms dev=%.1f ms" % (M, (t1 - t0) * 1e3, eng.timings()["predict"]), flush=True)
