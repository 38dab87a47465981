"""CPU (-m "not gpu"): the C-ABI library loads and exports every symbol include/gpb200.h declares,
the host-side mirror (kernel flattening, parameter plumbing, gradient order) behaves like the
reference's plumbing, the product fails loudly without a GPU, and the CPU-baseline port agrees with
the numpy oracle."""
import os
import re

import numpy as np
import pytest

import gpb200
from gpb200 import capi
from oracle import gp_oracle as orc
from conftest import make_data, kernel_zoo, ROOT


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "gpb200.h")).read()
    declared = sorted(set(re.findall(r"\b(gpb200_[a-z0-9_]+)\s*\(", hdr)) - {"gpb200_handle"})
    assert declared == capi.declared_symbols(), set(declared) ^ set(capi.declared_symbols())
    lib = gpb200.load_library()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.gpb200_version() >= 100


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback|no CUDA"):
        gpb200.Engine(0)


def test_flatten_programs():
    d = 4
    for name, k in kernel_zoo(d):
        ops, dims, theta, exposed = gpb200.flatten(k, d)
        assert ops.shape[1] == capi.OP_STRIDE and ops.shape[0] <= capi.MAX_OPS
        leaves = [o for o in ops if o[0] < 32]
        assert sum(o[2] for o in leaves) == theta.size
        assert [theta[i] for i in exposed] == pytest.approx(k.get_params())
        assert len(exposed) == k.num_params() == orc.num_params(k.spec())
        # post-order: a combinator's right child is the previous op
        depth = 0
        for o in ops:
            depth += 1 if o[0] < 32 else -1
            assert depth >= 1
        assert depth == 1


def test_masked_and_fixed_resolution():
    k = gpb200.Masked(gpb200.SEArd([0.1, 0.2], 0.3), [2, 0]) * gpb200.fix(gpb200.RQIso(0.4, 0.5, 0.6), "lσ")
    ops, dims, theta, exposed = gpb200.flatten(k, 3)
    assert list(dims[ops[0][3]:ops[0][3] + ops[0][4]]) == [2, 0]
    assert list(dims[ops[1][3]:ops[1][3] + ops[1][4]]) == [0, 1, 2]
    assert exposed == [0, 1, 2, 3, 5]
    k.set_params([1, 2, 3, 4, 5])
    assert gpb200.flatten(k, 3)[2].tolist() == [1, 2, 3, 4, 0.5, 5]
    with pytest.raises(ValueError):
        gpb200.flatten(gpb200.SEArd([0.1, 0.2], 0.3), 3)        # 2 length scales, 3 dims
    assert gpb200.fix(gpb200.fix(gpb200.SEIso(0.1, 0.2), "ll"), "lσ").num_params() == 0


class OracleEngine:
    """Stand-in with the Engine interface, backed by the CPU oracle -- TEST ONLY, to exercise the
    GPE host plumbing without a GPU."""

    def __init__(self):
        self.n_theta = 0

    def set_data(self, x):
        self.X = np.array(x); self.N, self.d = x.shape

    def set_kernel(self, ops, dims, n_theta):
        self.n_theta = n_theta

    def bind(self, gp):
        self.gp = gp

    def factorize(self, theta, ln, extra_nugget=0.0):
        self.ln = ln

    def mll(self, r):
        # unfixed spec: the device always sees the full parameter vector
        def unfix(s):
            if s[0] == "Fixed":
                return unfix(s[1])
            if s[0] in ("Sum", "Prod"):
                return (s[0], unfix(s[1]), unfix(s[2]))
            if s[0] == "Masked":
                return ("Masked", unfix(s[1]), s[2])
            return s
        self.spec = unfix(self.gp.kernel.spec())
        self.f = orc.mll_and_dmll(self.spec, self.X, r, self.ln)
        return self.f["alpha"], self.f["mll"]

    def grad_prepare(self):
        pass

    def grad_kernel(self, alpha=None):
        return self.f["dmll_kernel"], self.f["trA"]

    def predict(self, xs, alpha=None, want_var=True, full_cov=False):
        mu, v = orc.predict_f(self.spec, self.X, self.f, xs, full_cov=full_cov)
        return mu, (None if full_cov else np.diag(orc.cov(self.spec, xs, xs)) - (np.diag(orc.cov(self.spec, xs, xs)) - v)), (v if full_cov else None)


def _gp(kernel, mean, ln, X, y):
    eng = OracleEngine()
    gp = gpb200.GPE.__new__(gpb200.GPE)
    eng.bind(gp)
    gpb200.GPE.__init__(gp, X.T, y, mean, kernel, ln, engine=eng)
    return gp


def test_gpe_plumbing_gradient_order_and_flags():
    X, y, Xs = make_data(50, 2, 9, m=6)
    k = gpb200.SEIso(0.1, 0.2) + gpb200.fix(gpb200.RQIso(0.3, -0.5, 0.2), "lσ")
    gp = _gp(k, gpb200.MeanLin([0.1, -0.2]), -1.0, X, y)
    gp.update_target_and_dtarget()
    o = orc.mll_and_dmll(k.spec(), X, y, -1.0, ("MeanLin", [0.1, -0.2]))
    assert gp.mll == pytest.approx(o["mll"], rel=1e-12)
    assert gp.dmll.size == 1 + 2 + 4                          # [noise; mean; kernel(free)]  GPE.jl:298-324
    assert np.allclose(gp.dmll, o["dmll"], rtol=1e-10)
    assert gp.get_params().tolist() == pytest.approx([-1.0, 0.1, -0.2, 0.1, 0.2, 0.3, 0.2])
    gp.update_dmll(noise=False, domean=False)
    assert gp.dmll.size == 4
    gp.update_dmll(kern=False)
    assert gp.dmll.size == 3
    p = gp.get_params(); p[0] = -0.7; p[3] = 0.25
    gp.set_params(p)
    assert gp.logNoise == -0.7 and gp.kernel.get_params()[0] == 0.25
    with pytest.raises(ValueError):
        gp.set_params(p[:-1])
    with pytest.raises(ValueError):
        gp.predict_f(np.zeros((3, 4)))                        # GP.jl:65 ArgumentError
    with pytest.raises(ValueError):
        gpb200.GPE(X.T, y[:-1], gpb200.MeanZero(), k, -1.0, engine=OracleEngine())   # GPE.jl:41
    my, vy = gp.predict_y(Xs.T)
    mf, vf = gp.predict_f(Xs.T)
    assert np.allclose(vy, vf + np.exp(2 * gp.logNoise))


def test_cpu_baseline_port_matches_oracle():
    from oracle import cpu_baseline as cb
    X, y, _ = make_data(400, 8, 2)
    r = cb.seiso_mll_and_dmll(X, y, 0.3, 0.3, 0.3, 0.0)
    o = orc.mll_and_dmll(("SEIso", [0.3, 0.3]), X, y, 0.3, ("MeanConst", 0.0))
    assert abs(r["mll"] - o["mll"]) < 1e-10 * abs(o["mll"])
    assert np.allclose(r["dmll"], o["dmll"], rtol=1e-9)
    assert np.allclose(r["alpha"], o["alpha"], rtol=1e-9, atol=1e-12)


def test_gpe_optimize_exception_filter_and_rollback():
    """optimize! (src/optimize.jl:19-97): the target increases; a PosDefException / ArgumentError inside the
    objective yields Inf and rolls the parameters back instead of aborting (optimize.jl:46-61, 72-87)."""
    X, _, _ = make_data(40, 1, 21)
    y = np.sin(2 * X[:, 0])
    gp = _gp(gpb200.SEIso(0.5, 0.5), gpb200.MeanZero(), -1.0, X, y)
    gp.update_target_and_dtarget()
    t0 = gp.target
    calls = {"n": 0}
    real_factorize = gp._eng.factorize

    def flaky(theta, ln, extra_nugget=0.0):
        calls["n"] += 1
        if calls["n"] == 2:                       # second objective evaluation blows up like a failed cholesky!
            raise gpb200.PosDefException(7)
        return real_factorize(theta, ln, extra_nugget)

    gp._eng.factorize = flaky
    res = gp.optimize(maxiter=20)
    assert calls["n"] > 3                          # the optimiser carried on after the failure
    assert gp.target > t0
    assert np.all(np.isfinite(gp.get_params()))
    # kern=False keeps the kernel parameters fixed (test/optim.jl)
    k0 = gp.kernel.get_params()
    gp.optimize(kern=False, maxiter=5)
    assert gp.kernel.get_params() == k0


def test_gpe_push_and_fit_rebuild_state():
    """push!(gp, x, y) / fit!(gp, x, y) (src/GPE.jl:128-136, 530-539) replace the data wholesale."""
    X, y, _ = make_data(30, 2, 5)
    gp = _gp(gpb200.SEIso(0.1, 0.2), gpb200.MeanConst(0.0), -1.0, X[:20], y[:20])
    gp.push(X[20:].T, y[20:])
    assert gp.nobs == 30 and gp.alpha.size == 30
    o = orc.fit(gp.kernel.spec(), X, y, -1.0, ("MeanConst", 0.0))
    assert gp.mll == pytest.approx(o["mll"], rel=1e-12)
    with pytest.raises(ValueError):
        gp.push(np.zeros((3, 1)), [0.0])


class OracleEngineCV(OracleEngine):
    """OracleEngine plus the round-2 entry points the host mirror drives (cv_param, cv_block, inverse_diag, rand, append,
    set_option) -- backed by numpy, TEST ONLY: exercises the host-side assembly of crossvalidation.jl / GPEelastic.jl."""

    def __init__(self):
        super().__init__()
        self.options = {}
        self.appended = 0

    def set_option(self, key, value):
        self.options[key] = value

    def _Kinv(self):
        Ky = orc.gram(self.spec, self.X, self.ln)
        return np.linalg.inv(Ky)

    def inverse_diag(self):
        return np.diag(self._Kinv()).copy()

    def cv_param(self, param, alpha=None):
        Kinv = self._Kinv()
        a = self.f["alpha"] if alpha is None else alpha
        if param < 0:
            Zj = Kinv
        else:
            _, grads = orc.cov_and_grads(self.spec, self.X, None, want_grad=True)
            Zj = Kinv @ grads[param]
        self._M = Zj @ Kinv
        return Zj @ a, np.diag(self._M).copy()

    def cv_block(self, which, idx):
        M = self._Kinv() if which == 0 else self._M
        return M[np.ix_(idx, idx)].copy()

    def rand(self, xs, z, nugget=1e-10, alpha=None):
        mu, cov = orc.predict_f(self.spec, self.X, self.f, xs, full_cov=True)
        L = np.linalg.cholesky(cov + nugget * np.eye(xs.shape[0]))
        return mu, mu[None, :] + z @ L.T

    def append(self, xnew):
        if self.N + xnew.shape[0] > self.options.get("capacity", 0):
            raise ValueError("append: capacity exceeded")
        self.X = np.vstack([self.X, xnew]); self.N = self.X.shape[0]
        self.appended += 1


def _gp_cv(kernel, mean, ln, X, y, **kw):
    eng = OracleEngineCV()
    gp = gpb200.GPE.__new__(gpb200.GPE)
    eng.bind(gp)
    gpb200.GPE.__init__(gp, X.T, y, mean, kernel, ln, engine=eng, **kw)
    return gp, eng


def test_crossvalidation_host_assembly_matches_oracle():
    """dlogpdθ_LOO / predict_CVfold / logp_CVfold / dlogpdθ_CVfold (src/crossvalidation.jl:67-341): the O(N) / per-fold host
    assembly of gpe.py over the device's two vectors and sub-blocks equals the oracle's literal restatement."""
    X, y, _ = make_data(60, 2, 12)
    k = gpb200.SEIso(0.2, 0.1) + gpb200.fix(gpb200.Mat32Iso(0.4, -0.3), "lσ")
    gp, eng = _gp_cv(k, gpb200.MeanZero(), -0.6, X, y)
    gp.update_target_and_dtarget()
    f = orc.fit(k.spec(), X, y, -0.6)
    mu, s2 = gp.predict_LOO()
    mo, so = orc.predict_loo(f, y)
    assert np.allclose(mu, mo, rtol=1e-9) and np.allclose(s2, so, rtol=1e-9)
    assert gp.logp_LOO() == pytest.approx(orc.logp_loo(f, y), rel=1e-10)
    assert np.allclose(gp.dlogp_LOO(), orc.dlogp_loo(k.spec(), X, y, f, -0.6), rtol=1e-8, atol=1e-10)
    assert gp.dlogp_LOO(noise=False).size == 3 and gp.dlogp_LOO(kern=False).size == 1
    folds = [np.arange(0, 20), np.arange(20, 45), np.arange(45, 60)]
    mus, Sigs = gp.predict_CVfold(folds)
    muo, Sigo = orc.predict_cvfold(f, y, folds)
    assert all(np.allclose(a, b, rtol=1e-8) for a, b in zip(mus, muo)) and all(np.allclose(a, b, rtol=1e-8) for a, b in zip(Sigs, Sigo))
    assert gp.logp_CVfold(folds) == pytest.approx(orc.logp_cvfold(f, y, folds), rel=1e-9)
    assert np.allclose(gp.dlogp_CVfold(folds), orc.dlogp_cvfold(k.spec(), X, y, f, -0.6, folds), rtol=1e-7, atol=1e-9)
    gpm, _ = _gp_cv(k, gpb200.MeanConst(0.1), -0.6, X, y)
    with pytest.raises(NotImplementedError):                  # "I don't know how to do means yet" (crossvalidation.jl:168)
        gpm.dlogp_LOO(domean=True)


def test_elastic_append_and_rand_host_logic():
    """ElasticGPE append! (GPEelastic.jl:13-22): in-place extension inside the capacity, refit with `stepsize` more room beyond
    it, dimension / length checks; rand (GP.jl:120-146): shape (npred, n), mean function added to every draw."""
    X, y, Xs = make_data(40, 2, 3, m=5)
    k = gpb200.SEIso(0.1, 0.2)
    gp, eng = _gp_cv(k, gpb200.MeanConst(0.5), -1.0, X[:20], y[:20], capacity=30, stepsize=16)
    assert eng.options["capacity"] == 30
    gp.append(X[20:26].T, y[20:26])
    assert eng.appended == 1 and gp.nobs == 26 and gp.alpha.size == 26
    o = orc.fit(k.spec(), X[:26], y[:26], -1.0, ("MeanConst", 0.5))
    assert gp.mll == pytest.approx(o["mll"], rel=1e-12)       # update_target!(gp, kern=false, noise=false) on the longer y
    gp.append(X[26:40].T, y[26:40])                           # 40 > capacity 30: refit with 40 + stepsize reserved
    assert eng.appended == 1 and gp.nobs == 40 and eng.options["capacity"] == 40 + 16
    o = orc.fit(k.spec(), X, y, -1.0, ("MeanConst", 0.5))
    assert gp.mll == pytest.approx(o["mll"], rel=1e-12)
    with pytest.raises(ValueError):
        gp.append(np.zeros((3, 1)), [0.0])
    with pytest.raises(ValueError):
        gp.append(np.zeros((2, 3)), [0.0, 1.0])
    draws = gp.rand(Xs.T, 7, rng=np.random.default_rng(0))
    assert draws.shape == (5, 7)
    z = np.random.default_rng(0).standard_normal((7, 5))
    mu_f, cov = orc.predict_f(k.spec(), X, o, Xs, ("MeanConst", 0.5), full_cov=True)
    want = mu_f[None, :] + z @ np.linalg.cholesky(cov + 1e-10 * np.eye(5)).T
    assert np.allclose(draws.T, want, rtol=1e-9, atol=1e-10)
