"""Leaf timing probe: Cholesky phase at N with the column-per-barrier vs blocked diagonal-tile kernel.
python profiles/tools/probe_leaf.py [N]"""
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
for leaf in (1, 0, 1, 0):
    eng.set_option("leaf", leaf)
    for la in (1, 0):
        eng.set_option("lookahead", la)
        eng.factorize(theta, 0.3)
        alpha, mll = eng.mll(y)
        print("N=%d leaf=%d lookahead=%d: cholesky %.2f ms  mll %.10f" % (N, leaf, la, eng.timings()["cholesky"], mll), flush=True)
