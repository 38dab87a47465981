"""CPU oracle for the exact-GP hot path of GaussianProcesses.jl  --  TEST INFRASTRUCTURE ONLY.

This file is a numpy/LAPACK *restatement* of the reference's algorithm (Julia, /root/reference).
It is the checker for the CUDA path; it is never the thing shipped or measured as the product.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import it.

Parity pin: `tests/test_oracle_golden.py` checks this oracle against the one known-answer
vector the reference tree holds for this path (perf/benchmarks/simdata.csv + the output recorded
in perf/benchmarks/notebooks/benchmark_julia.ipynb cell 6), reproduced to 2e-12.

Reference lines restated (paths relative to /root/reference):
  src/kernels/distance.jl:41-104     direct-difference (weighted) squared Euclidean distances
  src/kernels/stationary.jl:25-66    cov_ij / dKij_dθ! dispatch through the metric
  src/kernels/se_iso.jl:39-50 … periodic.jl:45-51 etc.   leaf formulas (see `_LEAVES`)
  src/kernels/sum_kernel.jl:15-51, prod_kernel.jl:14-68  composite value + product rule
  src/kernels/masked_kernel.jl:23-63, fixed_kernel.jl:8-69
  src/GPE.jl:169-186                 update_cK!: K + exp(2 logNoise) I  (scalar or per-point)
  src/GP.jl:101-112                  make_posdef!: dpotrf('U')
  src/GPE.jl:202-212                 update_mll!: alpha = K_y \\ (y - mu); mll
  src/GPE.jl:151-164                 get_ααinvcKI!: potrs on -I, then ger  (A = αα' - K_y^-1)
  src/GPE.jl:219-241                 dmll_kern!: 1/2 sum_ij A_ij dK_ij/dθ
  src/GPE.jl:273-324                 dmll_noise / dmll_mean! / gradient order [noise; mean; kernel]
  src/GP.jl:25-79                    predictMVN!, predict_f (variance clamp at 0)

Kernel "spec" (neutral tuple form shared with the product's host mirror):
  ("SEIso", [ll, lσ])            ("SEArd", [ll_1..ll_d, lσ])
  ("Mat12Iso"|"Mat32Iso"|"Mat52Iso", [ll, lσ])     ("Mat12Ard"|..., [ll_1.., lσ])
  ("RQIso", [ll, lσ, lα])        ("RQArd", [ll_1.., lσ, lα])
  ("Periodic", [ll, lσ, lp])     ("LinIso", [ll])   ("LinArd", [ll_1..])
  ("Poly", [lc, lσ], deg)        ("Noise", [lσ])    ("Const", [lσ])
  ("Sum", left, right)           ("Prod", left, right)
  ("Masked", inner, [dims 0-based])                 ("Fixed", inner, [free idx 0-based])
x is stored point-major: shape (N, d)  (== Julia's d×N column-major).
"""
import math
import numpy as np

try:  # LAPACK through scipy's bundled OpenBLAS (the same backend family Julia ships)
    from scipy.linalg import lapack as _lapack
except Exception:  # pragma: no cover
    _lapack = None

LOG2PI = math.log(2.0 * math.pi)


# ----------------------------------------------------------------------------------------------
# distances  (src/kernels/distance.jl:41-104): sum_k (x_ki - x_kj)^2 [* w_k], direct differences
# ----------------------------------------------------------------------------------------------
def _sqdist(X1, X2, w=None):
    n1, d = X1.shape
    n2 = X2.shape[0]
    R = np.zeros((n1, n2))
    for k in range(d):  # same summation order as the reference's @simd loop over k
        diff = X1[:, k][:, None] - X2[:, k][None, :]
        if w is None:
            R += diff * diff
        else:
            R += diff * diff * w[k]
    return R


def _sqdist_k(X1, X2, k, wk=1.0):
    diff = X1[:, k][:, None] - X2[:, k][None, :]
    return diff * diff * wk


def _dot(X1, X2):
    n1, d = X1.shape
    S = np.zeros((n1, X2.shape[0]))
    for k in range(d):
        S += X1[:, k][:, None] * X2[:, k][None, :]
    return S


# ----------------------------------------------------------------------------------------------
# leaves: each returns (K, [dK/dθ_p ...])   (gradients only if want_grad)
# ----------------------------------------------------------------------------------------------
def _leaf(name, theta, extra, X1, X2, want_grad):
    th = [float(t) for t in theta]
    d = X1.shape[1]
    g = []
    if name == "SEIso":  # se_iso.jl:39-50
        l2, s2 = math.exp(2 * th[0]), math.exp(2 * th[1])
        r = _sqdist(X1, X2)
        K = s2 * np.exp(-0.5 * r / l2)
        if want_grad:
            g = [r / l2 * K, 2.0 * K]
    elif name == "SEArd":  # se_ard.jl:43-50
        il2 = np.exp(-2.0 * np.array(th[:-1]))
        s2 = math.exp(2 * th[-1])
        assert len(il2) == d, "SEArd: wrong number of length scales"
        r = _sqdist(X1, X2, il2)
        K = s2 * np.exp(-r / 2.0)
        if want_grad:
            g = [_sqdist_k(X1, X2, p, il2[p]) * K for p in range(d)] + [2.0 * K]
    elif name in ("Mat12Iso", "Mat32Iso", "Mat52Iso"):  # mat12_iso.jl:41-43, mat32_iso.jl:41-45, mat52_iso.jl:40-44
        l, s2 = math.exp(th[0]), math.exp(2 * th[1])
        r = np.sqrt(_sqdist(X1, X2))
        if name == "Mat12Iso":
            K = s2 * np.exp(-r / l)
            dll = r / l * K
        elif name == "Mat32Iso":
            s = math.sqrt(3.0) * r / l
            K = s2 * (1 + s) * np.exp(-s)
            dll = s2 * s ** 2 * np.exp(-s)
        else:
            s = math.sqrt(5.0) * r / l
            K = s2 * (1 + s + s ** 2 / 3.0) * np.exp(-s)
            dll = s2 / 3.0 * s ** 2 * (1 + s) * np.exp(-s)
        if want_grad:
            dll = np.where(r == 0.0, 0.0, dll)  # mat.jl:24-26
            g = [dll, 2.0 * K]
    elif name in ("Mat12Ard", "Mat32Ard", "Mat52Ard"):  # mat12_ard.jl:43-45, mat32_ard.jl:43-46, mat52_ard.jl:43-47
        il2 = np.exp(-2.0 * np.array(th[:-1]))
        s2 = math.exp(2 * th[-1])
        assert len(il2) == d
        r = np.sqrt(_sqdist(X1, X2, il2))
        if name == "Mat12Ard":
            K = s2 * np.exp(-r)
        elif name == "Mat32Ard":
            s = math.sqrt(3.0) * r
            K = s2 * (1 + s) * np.exp(-s)
        else:
            s = math.sqrt(5.0) * r
            K = s2 * (1 + s + s ** 2 / 3.0) * np.exp(-s)
        if want_grad:
            for p in range(d):
                wd = _sqdist_k(X1, X2, p, il2[p])
                with np.errstate(divide="ignore", invalid="ignore"):
                    if name == "Mat12Ard":
                        v = wd / r * K
                    elif name == "Mat32Ard":
                        v = 3.0 * s2 * wd * np.exp(-math.sqrt(3.0) * r)
                    else:
                        s = math.sqrt(5.0) * r
                        v = 5.0 / 3.0 * s2 * wd * (1 + s) * np.exp(-s)
                g.append(np.where(wd > 0, v, 0.0))  # mat.jl:5-18
            g.append(2.0 * K)
    elif name == "RQIso":  # rq_iso.jl:44-52
        l2, s2, al = math.exp(2 * th[0]), math.exp(2 * th[1]), math.exp(th[2])
        r = _sqdist(X1, X2)
        K = s2 * (1 + r / (2 * al * l2)) ** (-al)
        if want_grad:
            s = r / l2
            part = 1 + s / (2 * al)
            g = [s2 * s * part ** (-al - 1), 2.0 * K,
                 s2 * part ** (-al) * (s / (2 * part) - al * np.log(part))]
    elif name == "RQArd":  # rq_ard.jl:47-54
        il2 = np.exp(-2.0 * np.array(th[:-2]))
        s2, al = math.exp(2 * th[-2]), math.exp(th[-1])
        assert len(il2) == d
        r = _sqdist(X1, X2, il2)
        part = 1 + r / (2 * al)
        K = s2 * (1 + 0.5 * r / al) ** (-al)
        if want_grad:
            g = [s2 * _sqdist_k(X1, X2, p, il2[p]) * part ** (-al - 1) for p in range(d)]
            g.append(2.0 * K)
            g.append(s2 * part ** (-al) * (r / (2 * part) - al * np.log(part)))
    elif name == "Periodic":  # periodic.jl:45-51
        l2, s2, per = math.exp(2 * th[0]), math.exp(2 * th[1]), math.exp(th[2])
        r = np.sqrt(_sqdist(X1, X2))
        K = s2 * np.exp(-2.0 / l2 * np.sin(math.pi * r / per) ** 2)
        if want_grad:
            s = 2 * np.sin(math.pi * r / per) ** 2 / l2
            sp = math.pi * r / per
            t = 2.0 / l2
            g = [2 * s2 * s * np.exp(-s), 2.0 * K,
                 s2 * sp * t * np.sin(2 * sp) * np.exp(-t * np.sin(sp) ** 2)]
    elif name == "LinIso":  # lin_iso.jl:42,71
        l2 = math.exp(2 * th[0])
        K = _dot(X1, X2) / l2
        if want_grad:
            g = [-2.0 * K]
    elif name == "LinArd":  # lin_ard.jl:69-75,92
        l = np.exp(np.array(th))
        assert len(l) == d
        K = np.zeros((X1.shape[0], X2.shape[0]))
        parts = []
        for k in range(d):
            pk = X1[:, k][:, None] * X2[:, k][None, :] * (1.0 / l[k] ** 2)
            K += pk
            parts.append(pk)
        if want_grad:
            g = [-2.0 * pk for pk in parts]
    elif name == "Poly":  # poly.jl:44,69-70
        c, s2, deg = math.exp(th[0]), math.exp(2 * th[1]), int(extra)
        xy = _dot(X1, X2)
        K = s2 * (c + xy) ** deg
        if want_grad:
            g = [c * deg * s2 * (c + xy) ** (deg - 1), 2.0 * K]
    elif name == "Noise":  # noise.jl:31-52  (isapprox per coordinate, rtol = sqrt(eps))
        s2 = math.exp(2 * th[0])
        rtol = math.sqrt(np.finfo(float).eps)
        same = np.ones((X1.shape[0], X2.shape[0]), dtype=bool)
        for k in range(d):
            a = X1[:, k][:, None]
            b = X2[:, k][None, :]
            same &= np.abs(a - b) <= rtol * np.maximum(np.abs(a), np.abs(b))
        K = np.where(same, s2, 0.0)
        if want_grad:
            g = [2.0 * K]
    elif name == "Const":  # const.jl:41
        s2 = math.exp(2 * th[0])
        K = np.full((X1.shape[0], X2.shape[0]), s2)
        if want_grad:
            g = [2.0 * K]
    else:
        raise ValueError("unknown kernel leaf %r" % (name,))
    return K, g


def num_params(spec):
    name = spec[0]
    if name in ("Sum", "Prod"):
        return num_params(spec[1]) + num_params(spec[2])
    if name == "Masked":
        return num_params(spec[1])
    if name == "Fixed":
        return len(spec[2])
    return len(spec[1])


def cov_and_grads(spec, X1, X2=None, want_grad=False):
    """K(X1,X2) and the list of dK/dθ_p in get_params order (kernels.jl:31-71, 89-131)."""
    X1 = np.ascontiguousarray(X1, dtype=np.float64)
    X2 = X1 if X2 is None else np.ascontiguousarray(X2, dtype=np.float64)
    name = spec[0]
    if name in ("Sum", "Prod"):
        Kl, gl = cov_and_grads(spec[1], X1, X2, want_grad)
        Kr, gr = cov_and_grads(spec[2], X1, X2, want_grad)
        if name == "Sum":  # sum_kernel.jl:15-16, 34-51
            return Kl + Kr, gl + gr
        return Kl * Kr, [a * Kr for a in gl] + [Kl * b for b in gr]  # prod_kernel.jl:14-15, 39-68
    if name == "Masked":  # masked_kernel.jl:23,44-63
        dims = list(spec[2])
        return cov_and_grads(spec[1], X1[:, dims], X2[:, dims], want_grad)
    if name == "Fixed":  # fixed_kernel.jl:64-69
        K, g = cov_and_grads(spec[1], X1, X2, want_grad)
        return K, [g[f] for f in spec[2]] if want_grad else []
    extra = spec[2] if len(spec) > 2 else None
    return _leaf(name, spec[1], extra, X1, X2, want_grad)


def cov(spec, X1, X2=None):
    return cov_and_grads(spec, X1, X2, False)[0]


# ----------------------------------------------------------------------------------------------
# means (host-side in the product too; src/means/*.jl)
# ----------------------------------------------------------------------------------------------
def mean_and_grads(mspec, X):
    """mspec: ("MeanZero",) | ("MeanConst", beta) | ("MeanLin", [beta...]).  Returns mu[N], G[N,nm]."""
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    if mspec[0] == "MeanZero":
        return np.zeros(n), np.zeros((n, 0))
    if mspec[0] == "MeanConst":
        return np.full(n, float(mspec[1])), np.ones((n, 1))
    if mspec[0] == "MeanLin":
        b = np.asarray(mspec[1], dtype=np.float64)
        return X @ b, X.copy()
    raise ValueError(mspec)


# ----------------------------------------------------------------------------------------------
# the hot path
# ----------------------------------------------------------------------------------------------
def _potrf_upper(Ky):
    if _lapack is not None:
        U, info = _lapack.dpotrf(Ky, lower=0, clean=1, overwrite_a=0)
        if info != 0:
            raise np.linalg.LinAlgError("PosDefException(%d)" % info)
        return U
    return np.linalg.cholesky(Ky).T


def _potrs_upper(U, B):
    if _lapack is not None:
        X, info = _lapack.dpotrs(U, B, lower=0)
        assert info == 0
        return X
    import scipy.linalg as sl
    return sl.cho_solve((U, False), B)


def gram(spec, X, log_noise, extra_nugget=0.0):
    """K_y = K + exp(2 logNoise) I (scalar) or + diag(exp(2 logNoise_i)) (GPE.jl:169-186)."""
    K = cov(spec, X)
    ln = np.atleast_1d(np.asarray(log_noise, dtype=np.float64))
    if ln.size == 1:
        K = K + (math.exp(2 * float(ln[0])) + extra_nugget) * np.eye(X.shape[0])
    else:
        K = K + np.diag(np.exp(2 * ln) + extra_nugget)
    return K


def fit(spec, X, y, log_noise, mspec=("MeanZero",), extra_nugget=0.0):
    """update_mll! (GPE.jl:202-212).  Returns dict(U, alpha, mll, logdet, Ky)."""
    X = np.ascontiguousarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    Ky = gram(spec, X, log_noise, extra_nugget)
    U = _potrf_upper(Ky)
    mu, _ = mean_and_grads(mspec, X)
    r = y - mu
    alpha = _potrs_upper(U, r)
    logdet = 2.0 * np.sum(np.log(np.diag(U)))
    mll = -(r @ alpha + logdet + LOG2PI * X.shape[0]) / 2.0
    return dict(U=U, alpha=alpha, mll=mll, logdet=logdet, Ky=Ky, resid=r)


def mll_and_dmll(spec, X, y, log_noise, mspec=("MeanZero",), extra_nugget=0.0):
    """update_mll_and_dmll! (GPE.jl:332-335).  dmll order = [noise; mean; kernel] (GPE.jl:298-324)."""
    f = fit(spec, X, y, log_noise, mspec, extra_nugget)
    n = X.shape[0]
    A = _potrs_upper(f["U"], -np.eye(n))            # GPE.jl:157-162  (the reference's 2N^3 form)
    A += np.outer(f["alpha"], f["alpha"])           # GPE.jl:163  ger!
    _, grads = cov_and_grads(spec, X, None, True)
    dk = np.array([0.5 * np.sum(A * G) for G in grads])       # GPE.jl:226-239
    ln = np.atleast_1d(np.asarray(log_noise, dtype=np.float64))
    dnoise = math.exp(2 * float(ln[0])) * np.trace(A) if ln.size == 1 else float("nan")  # GPE.jl:273-275
    _, MG = mean_and_grads(mspec, X)
    dmean = MG.T @ f["alpha"]                       # GPE.jl:282-288
    f["dmll"] = np.concatenate([[dnoise], dmean, dk])
    f["dmll_kernel"] = dk
    f["trA"] = float(np.trace(A))
    f["A"] = A
    return f


def predict_f(spec, X, f, Xs, mspec=("MeanZero",), full_cov=False):
    """predict_f / predictMVN! (GP.jl:25-79).  f = result of fit()."""
    import scipy.linalg as sl
    Xs = np.ascontiguousarray(Xs, dtype=np.float64)
    Kc = cov(spec, X, Xs)                           # N x M   (GP.jl:44)
    Kp = cov(spec, Xs, Xs)                          # M x M   (GP.jl:45)
    mx, _ = mean_and_grads(mspec, Xs)
    mu = mx + Kc.T @ f["alpha"]                     # GP.jl:26
    V = sl.solve_triangular(f["U"], Kc, trans="T", lower=False)   # whiten!: U^-T Kc  (GP.jl:27)
    if full_cov:
        return mu, Kp - V.T @ V                     # GP.jl:51-54
    var = np.diag(Kp) - np.sum(V * V, axis=0)
    return mu, np.maximum(var, 0.0)                 # GP.jl:75


# ----------------------------------------------------------------------------------------------
# FITC sparse strategy (src/sparse/fully_indep_train_conditional.jl), restated with dense numpy
# ----------------------------------------------------------------------------------------------
def fitc_fit(spec, X, Xu, y, log_noise, mspec=("MeanZero",), mode="FITC"):
    """update_cK!(::FullyIndepPDMat) (fitc.jl:134-156) + `\\` (fitc.jl:33-36) + logdet (fitc.jl:77) + mll."""
    import scipy.linalg as sl
    X = np.ascontiguousarray(X, dtype=np.float64); Xu = np.ascontiguousarray(Xu, dtype=np.float64)
    n = X.shape[0]
    Kuu = cov(spec, Xu) + 1e-10 * np.eye(Xu.shape[0])          # make_posdef!(..., nugget=1e-10)  fitc.jl:140
    Uuu = _potrf_upper(Kuu)
    Kuf = cov(spec, Xu, X)                                      # fitc.jl:142
    Kdiag = np.array([cov(spec, X[i:i + 1], X[i:i + 1])[0, 0] for i in range(n)]) if n <= 64 else \
        np.diag(cov(spec, X[:1], X[:1])).repeat(n) if False else _kdiag(spec, X)
    W = sl.solve_triangular(Uuu, Kuf, trans="T", lower=False)   # U^-T Kuf ; invquad = column norms  fitc.jl:147
    Qdiag = np.sum(W * W, axis=0)
    Lam = math.exp(2 * log_noise) + Kdiag - Qdiag               # fitc.jl:148
    if mode != "FITC":                                          # SoR / DTC: Σ ≈ Qff + σ²I  (sor.jl:96-106)
        Lam = np.full(n, math.exp(2 * log_noise))
    SQR = Kuf @ (Kuf / Lam).T + Kuu + 1e-10 * np.eye(Xu.shape[0])     # fitc.jl:150-153
    Us = _potrf_upper(SQR)
    mu, MG = mean_and_grads(mspec, X)
    r = y - mu
    Lk = sl.solve_triangular(Us, Kuf, trans="T", lower=False)   # whiten(ΣQR, Kuf)
    alpha = (r - Lk.T @ (Lk @ (r / Lam))) / Lam                 # fitc.jl:33-36
    logdet = 2 * np.sum(np.log(np.diag(Us))) - 2 * np.sum(np.log(np.diag(Uuu))) + np.sum(np.log(Lam))   # fitc.jl:77
    mll = -(r @ alpha + logdet + LOG2PI * n) / 2.0
    LkL = Lk / Lam
    dnoise = math.exp(2 * log_noise) * (alpha @ alpha - np.sum(1.0 / Lam) + np.sum(LkL * LkL))   # fitc.jl:243-257
    alpha_u = _potrs_upper(Us, Kuf @ (r / Lam))                 # fitc.jl:279-286
    Sigma = (W.T @ W + np.diag(Lam)) if n <= 4096 else None      # Matrix(::FullyIndepPDMat)  fitc.jl:78-83
    return dict(mode=mode, alpha=alpha, mll=mll, logdet=logdet, Lam=Lam, Uuu=Uuu, Us=Us, alpha_u=alpha_u, dmll_noise=dnoise,
                dmll_mean=MG.T @ alpha, Sigma=Sigma, resid=r)


def _kdiag(spec, X):
    out = np.empty(X.shape[0])
    for i0 in range(0, X.shape[0], 512):
        blk = X[i0:i0 + 512]
        out[i0:i0 + 512] = np.diag(cov(spec, blk, blk))
    return out


def fitc_predict_full(spec, X, Xu, f, Xs, mspec=("MeanZero",)):
    """predictMVN with the full covariance: SoR  Σ = Lck'Lck, Lck = whiten(ΣQR, Kux)  (sor.jl:302-321);
    DTC / FITC  Σ = Σxx - Qxx + Σ_SoR  (dtc.jl:41-59, fitc.jl:324-332)."""
    import scipy.linalg as sl
    Kux = cov(spec, Xu, Xs)
    mx, _ = mean_and_grads(mspec, Xs)
    mu = mx + Kux.T @ f["alpha_u"]
    Lck = sl.solve_triangular(f["Us"], Kux, trans="T", lower=False)
    S_sor = Lck.T @ Lck
    if f.get("mode", "FITC") == "SoR":
        return mu, S_sor
    Lq = sl.solve_triangular(f["Uuu"], Kux, trans="T", lower=False)
    return mu, cov(spec, Xs, Xs) - Lq.T @ Lq + S_sor


def fitc_predict(spec, X, Xu, f, Xs, mspec=("MeanZero",)):
    """predictMVN (fitc.jl:324-332 -> dtc.jl:41-59 -> sor.jl:302-321): mu = mx + Kxu alpha_u;
    Sigma = Kxx - Qxx + Kxu ΣQR^-1 Kux (diagonal returned, clamped at 0 like GP.jl:75)."""
    import scipy.linalg as sl
    Kux = cov(spec, Xu, Xs)
    mx, _ = mean_and_grads(mspec, Xs)
    mu = mx + Kux.T @ f["alpha_u"]
    q = np.sum(sl.solve_triangular(f["Uuu"], Kux, trans="T", lower=False) ** 2, axis=0)
    s = np.sum(sl.solve_triangular(f["Us"], Kux, trans="T", lower=False) ** 2, axis=0)
    if f.get("mode", "FITC") == "SoR":                            # sor.jl:302-321: Σ = Kxu ΣQR⁻¹ Kux
        return mu, np.maximum(s, 0.0)
    return mu, np.maximum(_kdiag(spec, Xs) - q + s, 0.0)


def fitc_dmll_kern(spec, X, Xu, f):
    """Kernel-parameter gradient of the FITC mll, restating the reference literally:
    SoR part  dmll_kern!(..., ::SubsetOfRegsStrategy)   src/sparse/subsetofregressors.jl:219-253
              with precompute! (Kuu⁻¹Kuf, Kuu⁻¹KufΣ⁻¹y, Σ⁻¹Kfu)  subsetofregressors.jl:140-151
    FITC part dmll_kern!(..., ::FullyIndepStrat)         src/sparse/fully_indep_train_conditional.jl:200-234
    `f` = result of fitc_fit (dense N x M matrices: small sizes only)."""
    import scipy.linalg as sl
    X = np.ascontiguousarray(X, dtype=np.float64); Xu = np.ascontiguousarray(Xu, dtype=np.float64)
    n = X.shape[0]
    alpha, Lam, Uuu, Us = f["alpha"], f["Lam"], f["Uuu"], f["Us"]
    Kuf, gKuf = cov_and_grads(spec, Xu, X, True)
    _, gKuu = cov_and_grads(spec, Xu, None, True)
    gKdiag = [np.array([cov_and_grads(spec, X[i:i + 1], X[i:i + 1], True)[1][p][0, 0] for i in range(n)])
              for p in range(len(gKuu))]
    KuuinvKuf = _potrs_upper(Uuu, Kuf)                                   # Kuu \ Kuf            sor.jl:146
    b = _potrs_upper(Uuu, Kuf @ alpha)                                   # Kuu⁻¹KufΣ⁻¹y         sor.jl:147

    def solve_Sigma(B):                                                  # cK \ B  (fitc.jl:33-36)
        Lk = sl.solve_triangular(Us, Kuf, trans="T", lower=False)
        return (B - Lk.T @ (Lk @ (B / Lam[:, None]))) / Lam[:, None]

    SinvKfu = solve_Sigma(Kuf.T)                                         # Σ⁻¹Kfu               sor.jl:148
    out = []
    for p in range(len(gKuu)):
        dKuu, dKuf = gKuu[p], gKuf[p]
        V = 2 * alpha @ (dKuf.T @ b) - b @ (dKuu @ b)                    # sor.jl:246-247
        T = 2 * np.sum(solve_Sigma(dKuf.T) * KuuinvKuf.T)                # sor.jl:248
        T -= np.sum(SinvKfu.T * (_potrs_upper(Uuu, dKuu) @ KuuinvKuf))   # sor.jl:249
        g = (V - T) / 2.0
        if f.get("mode", "FITC") != "FITC":                              # SoR / DTC: no Λ-derivative part
            out.append(g)
            continue
        dLam = gKdiag[p] + np.sum(KuuinvKuf * (dKuu @ KuuinvKuf), axis=0) - 2 * np.sum(dKuf * KuuinvKuf, axis=0)   # fitc.jl:222-227
        V2 = alpha @ (dLam * alpha)                                      # fitc.jl:228
        Lsl = sl.solve_triangular(Us, Kuf / Lam, trans="T", lower=False)
        T2 = np.sum(dLam / Lam) - np.sum(Lsl * (Lsl * dLam))             # trinvAB  fitc.jl:63-67
        out.append(g + (V2 - T2) / 2.0)
    return np.array(out)


# ----------------------------------------------------------------------------------------------
# cross-validation (src/crossvalidation.jl) -- literal restatement with host matrices
# ----------------------------------------------------------------------------------------------
def _inv_sigma(f):
    n = f["U"].shape[0]
    return _potrs_upper(f["U"], np.eye(n))                       # inv(Σ), crossvalidation.jl:9


def predict_loo(f, y):
    """predict_LOO (crossvalidation.jl:8-13): sigma_i^2 = 1 / inv(Σ)_ii ; mu_i = y_i - alpha_i sigma_i^2"""
    invS = _inv_sigma(f)
    s2 = 1.0 / np.diag(invS)
    return -f["alpha"] * s2 + y, s2


def logp_loo(f, y):
    """logp_LOO (crossvalidation.jl:48-55): sum of Normal log-pdfs"""
    mu, s2 = predict_loo(f, y)
    return float(np.sum(-0.5 * np.log(2.0 * np.pi * s2) - 0.5 * (y - mu) ** 2 / s2))


def _loo_component(invS, alpha, y, Zj):
    """the per-parameter body of dlogpdθ_LOO_kern! / dlogpdσ2_LOO (crossvalidation.jl:86-108, 124-141), before the -1/2"""
    s2 = 1.0 / np.diag(invS)
    mu = -alpha * s2 + y
    ZjSinv = np.diag(Zj @ invS)
    ds2 = ZjSinv * s2 ** 2
    dmu = (Zj @ alpha) * s2 - alpha * ds2
    out = 0.0
    out -= np.sum(2.0 * (y - mu) / s2 * dmu)
    out -= np.sum((y - mu) ** 2 * ZjSinv)
    out += np.sum(ZjSinv * s2)
    return out


def dlogp_loo(spec, X, y, f, log_noise, noise=True, kern=True):
    """dlogpdθ_LOO (crossvalidation.jl:150-178): [noise; kernel] (mean parameters are not supported by the reference)"""
    invS = _inv_sigma(f)
    alpha = f["alpha"]
    out = []
    if noise:
        out.append(-_loo_component(invS, alpha, y, invS) / 2.0 * 2.0 * math.exp(2.0 * log_noise))
    if kern:
        _, grads = cov_and_grads(spec, X, None, want_grad=True)
        for dK in grads:
            out.append(-0.5 * _loo_component(invS, alpha, y, invS @ dK))
    return np.array(out)


def predict_cvfold(f, y, folds):
    """predict_CVfold (crossvalidation.jl:180-191)"""
    invS = _inv_sigma(f)
    mus, Sigs = [], []
    for V in folds:
        V = np.asarray(V)
        SVT = np.linalg.inv(invS[np.ix_(V, V)])
        mus.append(y[V] - SVT @ f["alpha"][V])
        Sigs.append(SVT)
    return mus, Sigs


def logp_cvfold(f, y, folds):
    """logp_CVfold (crossvalidation.jl:225-237): multivariate normal log-pdf per fold, nugget 1e-10"""
    mus, Sigs = predict_cvfold(f, y, folds)
    cv = 0.0
    for mu, S, V in zip(mus, Sigs, folds):
        V = np.asarray(V)
        S = S + 1e-10 * np.eye(len(V))
        L = np.linalg.cholesky(S)
        z = np.linalg.solve(L, y[V] - mu)
        cv += -0.5 * (z @ z) - np.sum(np.log(np.diag(L))) - 0.5 * len(V) * LOG2PI
    return float(cv)


def _fold_component(invS, alpha, ZjSinv, Zja, V):
    """gradient_fold (crossvalidation.jl:248-262)"""
    V = np.asarray(V)
    SVTinv = invS[np.ix_(V, V)]
    SVTa = np.linalg.solve(SVTinv, alpha[V])
    ZVV = ZjSinv[np.ix_(V, V)]
    out = -2.0 * (SVTa @ Zja[V])
    out += SVTa @ (ZVV @ SVTa)
    out += np.trace(np.linalg.solve(SVTinv, ZVV))
    return out


def dlogp_cvfold(spec, X, y, f, log_noise, folds, noise=True, kern=True):
    """dlogpdθ_CVfold (crossvalidation.jl:264-341): [noise; kernel]"""
    invS = _inv_sigma(f)
    alpha = f["alpha"]
    out = []
    if noise:
        Zj = invS
        comp = sum(_fold_component(invS, alpha, Zj @ invS, Zj @ alpha, V) for V in folds)
        out.append(-comp / 2.0 * 2.0 * math.exp(2.0 * log_noise))
    if kern:
        _, grads = cov_and_grads(spec, X, None, want_grad=True)
        for dK in grads:
            Zj = invS @ dK
            comp = sum(_fold_component(invS, alpha, Zj @ invS, Zj @ alpha, V) for V in folds)
            out.append(-0.5 * comp)
    return np.array(out)
