"""C3 phase probe (no oracle, seconds): Mauna-Loa-style composite kernel, N=16384, d=1, 13 parameters.
python profiles/tools/probe_c3.py  -> phase timers of one update_mll_and_dmll!, plus a directional finite-difference check of
the gradient against the device mll (the oracle comparison at this size lives in tests/test_gpu_configs.py)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "gaussianprocesses.jl_b200"))
import gpb200 as g
from gpb200 import capi
if os.environ.get("GPB200_PROBE_LIB"):                       # A/B against a library built from another commit
    capi.LIB_PATH = os.environ["GPB200_PROBE_LIB"]
    print("library:", capi.LIB_PATH, flush=True)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
rng = np.random.default_rng(3)
x = np.sort(rng.uniform(1958, 2004, N))
y = 315 + 1.5 * (x - 1958) + 3 * np.sin(2 * np.pi * x) + 0.3 * rng.standard_normal(N)
y = (y - y.mean()) / y.std()
k = g.SEArd([4.0], 0.0) + g.Periodic(0.0, 0.0, 0.0) * g.SEArd([4.0], 0.0) + g.RQIso(0.0, 0.0, -1.0) + g.SEArd([-2.0], -2.0)
gp = g.GPE(x[None, :], y, g.MeanZero(), k, 0.0)
for it in range(3):
    gp.update_target_and_dtarget()
    tm = gp._eng.timings()
    print("C3 N=%d eval %d: " % (N, it) + ", ".join("%s %.3f" % (kk, tm[kk]) for kk in ("gram", "cholesky", "solve_mll", "inverse", "trace")) +
          " ms | mll %.12f" % gp.mll, flush=True)
print("dmll", np.array2string(gp.dmll, precision=10), flush=True)
p0 = gp.get_params(); g0 = gp.dtarget.copy()
dirv = np.random.default_rng(0).standard_normal(p0.size); dirv /= np.linalg.norm(dirv)
h = 1e-5
gp.set_params(p0 + h * dirv); gp.update_target(); tp = gp.target
gp.set_params(p0 - h * dirv); gp.update_target(); tmn = gp.target
fd = (tp - tmn) / (2 * h)
print("directional derivative: fd %.9f analytic %.9f rel %.2e" % (fd, g0 @ dirv, abs(fd - g0 @ dirv) / max(abs(fd), 1e-300)), flush=True)
