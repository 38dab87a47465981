"""CPU baseline = the reference's algorithm for update_mll_and_dmll! on the host cores
(TEST / BENCH INFRASTRUCTURE).  Julia is not installed in this image, so this is the oracle "port":
  scalar single-threaded cov! loop          (ref_loops.c, as src/kernels/kernels.jl:39-50)
  K + exp(2 logNoise) I ; dpotrf('U')        (src/GPE.jl:173-174, src/GP.jl:104-110)   OpenBLAS, all cores
  alpha = dpotrs(y - mu) ; mll               (src/GPE.jl:206-210)
  A = dpotrs(-I) ; dger(alpha, alpha)        (src/GPE.jl:157-163)  -- the reference's 2N^3 formulation, kept
  scalar single-threaded dmll_kern! loop     (ref_loops.c, as src/GPE.jl:219-241)
Returns per-phase seconds and the results (checked against the numpy oracle in tests)."""
import ctypes as C
import math
import os
import subprocess
import time

import numpy as np
from scipy.linalg import lapack

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libgporacle.so")
_lib = None


def build():
    if not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "ref_loops.c")):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        dp = C.POINTER(C.c_double)
        _lib.cov_seiso_sym.argtypes = [dp, dp, C.c_int, C.c_long, C.c_double, C.c_double]
        _lib.dmll_kern_seiso.argtypes = [dp, dp, dp, C.c_int, C.c_long, C.c_double, C.c_double]
    return _lib


def host_threads():
    try:
        from threadpoolctl import threadpool_info
        return max([p.get("num_threads", 1) for p in threadpool_info()] or [os.cpu_count() or 1])
    except Exception:
        return os.cpu_count() or 1


def seiso_mll_and_dmll(x_pm, y, ll, lsig, log_noise, mean_const=0.0):
    """x_pm (N, d).  Returns dict(mll, dmll=[noise, beta, ll, lσ], alpha, seconds={...})."""
    lib = _load()
    dp = C.POINTER(C.c_double)
    x = np.ascontiguousarray(x_pm, dtype=np.float64)
    N, d = x.shape
    l2, s2 = math.exp(2 * ll), math.exp(2 * lsig)
    sec = {}
    t = time.perf_counter()
    K = np.empty((N, N), order="F")
    lib.cov_seiso_sym(K.ctypes.data_as(dp), x.ctypes.data_as(dp), d, N, l2, s2)
    sec["cov_loop"] = time.perf_counter() - t
    t = time.perf_counter()
    K[np.diag_indices(N)] += math.exp(2 * log_noise)
    U, info = lapack.dpotrf(K, lower=0, clean=0, overwrite_a=1)
    if info != 0:
        raise np.linalg.LinAlgError("PosDefException(%d)" % info)
    sec["dpotrf"] = time.perf_counter() - t
    t = time.perf_counter()
    r = y - mean_const
    alpha, _ = lapack.dpotrs(U, r, lower=0)
    logdet = 2.0 * np.sum(np.log(np.diag(U)))
    mll = -(r @ alpha + logdet + math.log(2 * math.pi) * N) / 2.0
    sec["solve_mll"] = time.perf_counter() - t
    t = time.perf_counter()
    A = np.zeros((N, N), order="F")
    A[np.diag_indices(N)] = -1.0
    A, _ = lapack.dpotrs(U, A, lower=0, overwrite_b=1)
    A += np.outer(alpha, alpha)
    sec["potrs_identity_ger"] = time.perf_counter() - t
    t = time.perf_counter()
    gk = np.zeros(2)
    A = np.asfortranarray(A)
    lib.dmll_kern_seiso(gk.ctypes.data_as(dp), A.ctypes.data_as(dp), x.ctypes.data_as(dp), d, N, l2, s2)
    dnoise = math.exp(2 * log_noise) * np.trace(A)
    dmean = np.sum(alpha)
    sec["dmll_loop"] = time.perf_counter() - t
    sec["total"] = sum(sec.values())
    return dict(mll=mll, dmll=np.array([dnoise, dmean, gk[0], gk[1]]), alpha=alpha, seconds=sec)
