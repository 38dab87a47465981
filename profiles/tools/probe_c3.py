"""Config C3 of SURVEY.md §8(d): Mauna-Loa-style composite kernel, N=16384, d=1 (generic kernel-program path)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "gaussianprocesses.jl_b200"))
import gpb200 as g
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
rng = np.random.default_rng(3)
x = np.sort(rng.uniform(1958, 2004, N))
y = 315 + 1.5 * (x - 1958) + 3 * np.sin(2 * np.pi * x) + 0.3 * rng.standard_normal(N)
y = (y - y.mean()) / y.std()
k = g.SEArd([4.0], 0.0) + g.Periodic(0.0, 0.0, 0.0) * g.SEArd([4.0], 0.0) + g.RQIso(0.0, 0.0, -1.0) + g.SEArd([-2.0], -2.0)
gp = g.GPE(x[None, :], y, g.MeanZero(), k, 0.0)
for rep in range(3):
    t0 = time.time(); gp.update_target_and_dtarget(); t1 = time.time()
    print("C3 N=%d rep=%d wall=%.1f ms mll=%.6f" % (N, rep, (t1 - t0) * 1e3, gp.mll), {k_: round(v, 2) for k_, v in gp._eng.timings().items()}, flush=True)
print("dmll", gp.dmll)
xs = rng.uniform(2004, 2024, 2048)
t0 = time.time(); mu, s2 = gp.predict_f(xs[None, :]); print("predict 2048: %.1f ms" % ((time.time() - t0) * 1e3))
