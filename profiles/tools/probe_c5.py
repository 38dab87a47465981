"""Config C5 of SURVEY.md §8(d): FITC SEIso(log 4, 0), N=1e6, M=8192 inducing, d=32, FP64, 1 GPU."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "gaussianprocesses.jl_b200"))
import gpb200 as g
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
M = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
d = 32
rng = np.random.default_rng(5)
X = rng.standard_normal((N, d))
y = np.sin(X.sum(axis=1) / np.sqrt(d)) + 0.1 * rng.standard_normal(N)
Xu = X[rng.permutation(N)[:M]]
t0 = time.time()
gp = g.FITC(X.T, Xu.T, y, g.MeanConst(float(y.mean())), g.SEIso(np.log(4.0), 0.0), np.log(0.1))
print("C5 construct (upload + update_cK! + mll): %.2f s  mll=%.6f" % (time.time() - t0, gp.mll), flush=True)
for rep in range(2):
    t0 = time.time(); gp.update_mll(); t1 = time.time()
    fl = 2.0 * M * M * N + 2.0 * M ** 3 / 3
    print("C5 N=%d M=%d update_mll: %.2f s  (%.1f TFLOP/s on 2M^2N+2M^3/3)  mll=%.6f" % (N, M, t1 - t0, fl / (t1 - t0) * 1e-12, gp.mll), flush=True)
t0 = time.time(); gp.update_dmll_noise_mean(); print("noise+mean gradient: %.2f s" % (time.time() - t0), gp.dmll, flush=True)
Xs = rng.standard_normal((4096, d))
t0 = time.time(); mu, s2 = gp.predict_f(Xs.T); print("predict_f 4096: %.3f s" % (time.time() - t0), mu[:3], s2[:3], flush=True)
