"""CPU: pin the oracle against the only known-answer vector the reference tree holds for this path
(perf/benchmarks/simdata.csv + benchmark_julia.ipynb cell 6), and check its
analytic gradients against finite differences the way test/kernels.jl:148-164 does."""
import math
import os

import numpy as np
import pytest

from oracle import gp_oracle as orc
from conftest import make_data, kernel_zoo

GOLD = os.path.join(os.path.dirname(__file__), "golden", "simdata_kat.npz")


@pytest.fixture(scope="module")
def kat():
    d = np.load(GOLD)
    return d["X"], d["Y"], float(d["recorded_mll"]), d["recorded_dmll"]


def test_oracle_reproduces_recorded_reference_output(kat):
    X, Y, mll_rec, dmll_rec = kat
    # the 2018 run carried a 1e-5 jitter on the diagonal
    f = orc.mll_and_dmll(("SEIso", [0.0, 0.0]), X, Y, 0.0, ("MeanConst", 0.0), extra_nugget=1e-5)
    assert abs(f["mll"] - mll_rec) < 1e-10 * abs(mll_rec)
    # dmll was printed to 6 significant figures in the notebook
    assert np.allclose(f["dmll"], dmll_rec, rtol=2e-6, atol=0)


def test_oracle_current_source_semantics(kat):
    X, Y, _, _ = kat
    f = orc.mll_and_dmll(("SEIso", [0.0, 0.0]), X, Y, 0.0, ("MeanConst", 0.0))
    assert abs(f["mll"] - (-4536.25646128023)) < 1e-8
    ref = np.array([-689.6318902132696, -15.731253155037956, 71.19489031472071, -667.2676571953316])
    assert np.allclose(f["dmll"], ref, rtol=1e-9)
    assert abs(np.abs(f["alpha"]).sum() - 1344.600310514208) < 1e-7


@pytest.mark.parametrize("idx", range(20))
def test_oracle_gradient_vs_finite_difference(idx):
    d = 3
    X, y, _ = make_data(40, d, 11)
    name, k = kernel_zoo(d)[idx]
    spec = k.spec()
    if name == "LinArd+Noise":
        pytest.skip("Noise kernel is discontinuous in x only; fine -- covered by value tests")
    ln = -0.3
    f = orc.mll_and_dmll(spec, X, y, ln, ("MeanConst", 0.2))
    hyp0 = np.array(k.get_params())
    g_fd = np.zeros_like(hyp0)
    for p in range(hyp0.size):
        for sgn in (+1, -1):
            h = hyp0.copy(); h[p] += sgn * 1e-5
            k.set_params(list(h))
            g_fd[p] += sgn * orc.fit(k.spec(), X, y, ln, ("MeanConst", 0.2))["mll"]
        g_fd[p] /= 2e-5
    k.set_params(list(hyp0))
    got = f["dmll"][2:]
    assert np.allclose(got, g_fd, rtol=1e-5, atol=1e-6), (name, got, g_fd)
    # noise and mean parts
    fn = lambda l, b: orc.fit(spec, X, y, l, ("MeanConst", b))["mll"]
    assert abs(f["dmll"][0] - (fn(ln + 1e-5, 0.2) - fn(ln - 1e-5, 0.2)) / 2e-5) < 1e-5 * (1 + abs(f["dmll"][0]))
    assert abs(f["dmll"][1] - (fn(ln, 0.2 + 1e-5) - fn(ln, 0.2 - 1e-5)) / 2e-5) < 1e-5 * (1 + abs(f["dmll"][1]))


def test_oracle_predict_interpolates():
    # test/gp.jl:32-38: predictive mean at the training inputs ~ y (atol 0.1), diag(full_cov) == var
    X, _, _ = make_data(60, 2, 5)
    y = np.sin(X.sum(axis=1))
    spec = ("SEIso", [0.0, 0.0])
    f = orc.fit(spec, X, y, -3.0)
    mu, var = orc.predict_f(spec, X, f, X)
    assert np.max(np.abs(mu - y)) < 0.1
    mu2, cov = orc.predict_f(spec, X, f, X[:7], full_cov=True)
    assert np.allclose(np.diag(cov), var[:7], atol=1e-10)


def test_oracle_fitc_matches_dense_formulation():
    """The reference's RNG-free FITC checks (test/test_sparse.jl:121-144): sparse mll / alpha / logdet equal
    those of a dense GP built from Matrix(cK); analytic noise gradient equals finite differences."""
    import scipy.linalg as sl
    rng = np.random.default_rng(0)
    N, M, d = 400, 30, 2
    X = rng.standard_normal((N, d)); y = np.sin(X.sum(1)) + 0.1 * rng.standard_normal(N)
    Xu = X[rng.permutation(N)[:M]]
    spec = ("SEIso", [-0.7, 0.1])
    f = orc.fitc_fit(spec, X, Xu, y, -1.0, ("MeanConst", 0.1))
    c = sl.cho_factor(f["Sigma"]); a = sl.cho_solve(c, f["resid"]); ld = 2 * np.sum(np.log(np.diag(c[0])))
    dense_mll = -(f["resid"] @ a + ld + orc.LOG2PI * N) / 2
    assert abs(f["mll"] - dense_mll) < 1e-6                    # test_sparse.jl:127
    assert abs(f["logdet"] - ld) < 1e-6                        # :130
    assert np.max(np.abs(a - f["alpha"])) < 1e-6               # :132
    e = 1e-5
    fd = (orc.fitc_fit(spec, X, Xu, y, -1.0 + e, ("MeanConst", 0.1))["mll"]
          - orc.fitc_fit(spec, X, Xu, y, -1.0 - e, ("MeanConst", 0.1))["mll"]) / (2 * e)
    assert abs(fd - f["dmll_noise"]) < 1e-3 * abs(fd)          # :134-144 (atol 1e-3 there)


@pytest.mark.parametrize("mode", ["FITC", "SoR", "DTC"])
def test_oracle_sparse_gradients_vs_finite_differences(mode):
    """test/test_sparse.jl:134-144 for every sparse strategy: the literal restatement of dmll_kern! / dmll_noise
    against finite differences of the sparse mll."""
    rng = np.random.default_rng(1)
    N, M, d = 250, 20, 2
    X = rng.standard_normal((N, d)); y = np.sin(X.sum(1)) + 0.1 * rng.standard_normal(N)
    Xu = X[rng.permutation(N)[:M]]
    th = [-0.6, 0.1, -0.4, -0.3, 0.2]
    mk = lambda t: ("Sum", ("SEIso", list(t[:2])), ("RQIso", list(t[2:])))
    f = orc.fitc_fit(mk(th), X, Xu, y, -1.0, mode=mode)
    g = orc.fitc_dmll_kern(mk(th), X, Xu, f)
    e = 1e-5
    for p in range(len(th)):
        tp, tm = list(th), list(th)
        tp[p] += e; tm[p] -= e
        fd = (orc.fitc_fit(mk(tp), X, Xu, y, -1.0, mode=mode)["mll"] - orc.fitc_fit(mk(tm), X, Xu, y, -1.0, mode=mode)["mll"]) / (2 * e)
        assert abs(fd - g[p]) < 1e-4 * (1 + abs(fd)), (mode, p, fd, g[p])
    fdn = (orc.fitc_fit(mk(th), X, Xu, y, -1.0 + e, mode=mode)["mll"] - orc.fitc_fit(mk(th), X, Xu, y, -1.0 - e, mode=mode)["mll"]) / (2 * e)
    assert abs(fdn - f["dmll_noise"]) < 1e-4 * (1 + abs(fdn))


def test_oracle_crossvalidation_matches_refits_and_finite_differences():
    """Pins the oracle's restatement of src/crossvalidation.jl the way the reference's own test/test_crossvalidation.jl does:
    analytic LOO / fold predictions == refitting without the held-out points; gradients == finite differences."""
    rng = np.random.default_rng(1)
    n = 24
    x = np.sort(rng.uniform(-2, 2, n))[:, None]
    y = np.abs(x[:, 0] - 5) * np.cos(2 * x[:, 0]) + 0.8 * rng.standard_normal(n)
    spec, ln = ("SEIso", [0.5, 0.8]), math.log(0.8)
    f = orc.fit(spec, x, y, ln)
    mu, s2 = orc.predict_loo(f, y)
    for i in (0, 7, 23):
        keep = np.arange(n) != i
        fi = orc.fit(spec, x[keep], y[keep], ln)
        m, v = orc.predict_f(spec, x[keep], fi, x[i:i + 1])
        assert abs(m[0] - mu[i]) < 1e-9 and abs(v[0] + math.exp(2 * ln) - s2[i]) < 1e-9
    folds = [np.arange(0, 5), np.arange(5, 14), np.arange(14, 24)]
    mus, Sigs = orc.predict_cvfold(f, y, folds)
    V = folds[1]
    keep = np.ones(n, bool); keep[V] = False
    fV = orc.fit(spec, x[keep], y[keep], ln)
    m, C = orc.predict_f(spec, x[keep], fV, x[V], full_cov=True)
    assert np.allclose(m, mus[1], atol=1e-9) and np.allclose(C + math.exp(2 * ln) * np.eye(V.size), Sigs[1], atol=1e-9)

    def crit(theta, lnn, which):
        sp = ("SEIso", list(theta))
        ff = orc.fit(sp, x, y, lnn)
        return orc.logp_loo(ff, y) if which == "loo" else orc.logp_cvfold(ff, y, folds)

    for which in ("loo", "fold"):
        g = orc.dlogp_loo(spec, x, y, f, ln) if which == "loo" else orc.dlogp_cvfold(spec, x, y, f, ln, folds)
        h = 1e-6
        fd = [(crit([0.5, 0.8], ln + h, which) - crit([0.5, 0.8], ln - h, which)) / (2 * h)]
        for p in range(2):
            tp = [0.5, 0.8]; tm = [0.5, 0.8]
            tp[p] += h; tm[p] -= h
            fd.append((crit(tp, ln, which) - crit(tm, ln, which)) / (2 * h))
        assert np.allclose(g, fd, rtol=1e-5, atol=1e-6), (which, g, fd)
