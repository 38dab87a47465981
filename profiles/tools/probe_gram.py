"""Gram / trace kernel probe at C2 size without the O(N^3) phases: python profiles/tools/probe_gram.py [N] [fast]
(get_gram-free: factorize would run the Cholesky, so the kernels are timed through gpb200's own phase timers on a
small-N factorization-free path: set_option("nb") is irrelevant; we call factorize once and read ms[gram], and grad_kernel
after grad_prepare for ms[trace])."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "gaussianprocesses.jl_b200"))
import gpb200
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
d = 8
rng = np.random.default_rng(1)
X = rng.standard_normal((N, d)); y = rng.standard_normal(N)
eng = gpb200.Engine(0)
eng.set_data(X)
ops, dims, theta, _ = gpb200.flatten(gpb200.SEIso(0.3, 0.3), d)
eng.set_kernel(ops, dims, theta.size)
res = {}
for fast in (1, 0, 1):
    eng.set_option("gram_fast", fast)
    eng.factorize(theta, 0.3)
    alpha, mll = eng.mll(y)
    eng.grad_prepare()
    gk, trA = eng.grad_kernel()
    tm = eng.timings()
    T = (N + 127) // 128
    tri = 8.0 * 128 * 128 * (T * (T + 1) // 2)
    print("N=%d gram_fast=%d: gram %.3f ms = %.0f GB/s, trace %.3f ms = %.0f GB/s | mll %.12f dmll %s trA %.10f"
          % (N, fast, tm["gram"], tri / tm["gram"] * 1e-6, tm["trace"], tri / tm["trace"] * 1e-6, mll, gk, trA), flush=True)
