"""One FULL-SIZE (N=32768, d=8) evaluation of the reference's CPU algorithm (oracle port, all host cores): the record
that validates bench.py's phase-wise extrapolation from the bounded sample.  ~5 min, ~26 GB RAM.
Writes gpurun_out/r02_cpu_full_size.json (copied to profiles/ when kept)."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
    os.environ[k] = str(os.cpu_count() or 1)
import numpy as np
import bench
from oracle import cpu_baseline as cb

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
cores = bench.pin_blas_threads()
X, y = bench.synth(N, bench.D, seed=1)
t0 = time.time()
r = cb.seiso_mll_and_dmll(X, y, bench.LL, bench.LSIG, bench.LNOISE, 0.0)
out = {"N": N, "d": bench.D, "cores": cores, "seconds": r["seconds"], "gflops_alg": bench.falg(N) / r["seconds"]["total"] * 1e-9,
       "mll": float(r["mll"]), "dmll": [float(v) for v in r["dmll"]], "alpha_l1": float(np.sum(np.abs(r["alpha"]))),
       "wall_s": time.time() - t0,
       "note": "scalar cov!/dmll_kern! loops (C, 1 thread) + dpotrf + dpotrs(-I) + dger (OpenBLAS, all cores); "
               "F_alg = N^3 + 2N^2; run concurrently with the GPU test-suite on the same box (CPU lightly shared)"}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "r02_cpu_full_size.json"), "w"), indent=1)
print(json.dumps(out))
