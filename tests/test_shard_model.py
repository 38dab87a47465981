"""CPU: numpy model of the ROW-SHARDED multi-GPU schedules (csrc/shard_impl.cuh) -- Cholesky with all-gathered panels,
fan-out triangular solves on L rows / U = L' rows, the backward panel sweep that produces K^-1 (TRTRI + LAUUM fused), and
the sharded predictive solve.  Every "rank" holds full-size arrays whose NON-owned rows are NaN, so any read outside the
rows a rank owns (or outside the panels it received) poisons the result; every product is written in the NT form
(C = A B') of the device GEMM, with the operand frames the C++ code uses.  The model is the specification the CUDA
schedules in shard_impl.cuh follow step by step."""
import numpy as np
import pytest

T = 4                      # model tile (128 on the device)


def nt(A, B):
    return A @ B.T


class Model:
    def __init__(self, K, R, rb):
        self.R, self.rb = R, rb
        self.Np = K.shape[0]
        assert self.Np % T == 0
        self.NB = rb * T
        self.nblk = (self.Np + self.NB - 1) // self.NB
        self.K = K
        self.F = [np.full_like(K, np.nan) for _ in range(R)]
        self.G = [np.full_like(K, np.nan) for _ in range(R)]
        for q in range(R):
            for r in self.own_rows(q):
                self.F[q][r, :] = 0.0
                self.G[q][r, :] = 0.0
                self.G[q][r, :r + 1] = K[r, :r + 1]              # Gram: own rows, lower part only
        self.Dinv = [np.zeros((self.Np, T)) for _ in range(R)]       # replicated small state
        self.DinvT = [np.zeros((self.Np, T)) for _ in range(R)]
        self.logd = [np.zeros(self.Np) for _ in range(R)]

    def owner_tile(self, t):
        return (t // self.rb) % self.R

    def own_rows(self, q, lo=0, hi=None):
        hi = self.Np if hi is None else hi
        return [r for r in range(lo, hi) if self.owner_tile(r // T) == q]

    def blk(self, k):
        c0 = k * self.NB
        return c0, min(self.NB, self.Np - c0), k % self.R

    # ------------------------------------------------------------------ Cholesky
    def cholesky(self):
        R = self.R
        for k in range(self.nblk):
            c0, nb, o = self.blk(k)
            r1 = c0 + nb
            # (a) owner factors its diagonal block in place (chol_panel limited to the block rows)
            Go = self.G[o]
            Lkk = np.linalg.cholesky(np.tril(Go[c0:r1, c0:r1]) + np.tril(Go[c0:r1, c0:r1], -1).T)
            self.F[o][c0:r1, c0:r1] = np.tril(Lkk) + np.triu(self.F[o][c0:r1, c0:r1], 1)
            stage_D = np.zeros((nb, T)); stage_DT = np.zeros((nb, T)); stage_ld = np.zeros(nb)
            for t in range(nb // T):
                W = np.linalg.inv(Lkk[t * T:(t + 1) * T, t * T:(t + 1) * T])
                stage_D[t * T:(t + 1) * T] = W
                stage_DT[t * T:(t + 1) * T] = W.T
            stage_ld[:] = 2.0 * np.log(np.diag(Lkk))
            stage_L = np.tril(Lkk)
            # (b) broadcast of the bundle; every rank unpacks it
            P = [np.full((self.Np, self.NB), np.nan) for _ in range(R)]
            for q in range(R):
                P[q][c0:r1, :nb] = stage_L
                self.Dinv[q][c0:r1] = stage_D; self.DinvT[q][c0:r1] = stage_DT; self.logd[q][c0:r1] = stage_ld
            # (c) every rank: TRSM of its own rows below the block (recursive, leaves through the inverted tiles)
            S = [dict() for _ in range(R)]
            for q in range(R):
                rows = self.own_rows(q, r1)
                if rows:
                    self.trsm_rows(q, rows, c0, nb, P[q])
                    X = self.G[q][rows, c0:r1]
                    self.F[q][rows, c0:r1] = X                 # scatter: permanent L rows ...
                    for r, xr in zip(rows, X):
                        S[q][r] = xr                           # ... and the all-gather contribution
            # (d)+(e) all-gather, unpacked into global row order
            for q in range(R):
                for p in range(R):
                    for r, xr in S[p].items():
                        P[q][r, :nb] = xr
            # owner keeps U = L' rows of its block (upper part of F): F[c0+i, r] = L[r, c0+i], r >= r1
            self.F[o][c0:r1, r1:] = P[o][r1:, :nb].T
            # (f) trailing update of own rows (lower trapezoid, NT GEMM with K = nb, row filter)
            for q in range(R):
                for i in self.own_rows(q, r1):
                    self.G[q][i, r1:i + 1] -= nt(P[q][i:i + 1, :nb], P[q][r1:i + 1, :nb])[0]

    def trsm_rows(self, q, rows, c0, nb, P):
        """G[rows, c0:c0+nb] <- G[rows, c0:c0+nb] L_kk^-T ; L_kk = P[c0:c0+nb, :nb], leaves use Dinv tiles"""
        G = self.G[q]

        def rec(off, n):
            if n == T:
                W = self.Dinv[q][c0 + off:c0 + off + T]                        # W_tt (lower)
                G[np.ix_(rows, range(c0 + off, c0 + off + T))] = nt(G[np.ix_(rows, range(c0 + off, c0 + off + T))], W)
                return
            n1 = T
            while n1 * 2 < n:
                n1 *= 2
            n2 = n - n1
            rec(off, n1)
            A = G[np.ix_(rows, range(c0 + off, c0 + off + n1))]
            B = P[c0 + off + n1:c0 + off + n, off:off + n1]                    # L21
            G[np.ix_(rows, range(c0 + off + n1, c0 + off + n))] -= nt(A, B)
            rec(off + n1, n2)
        rec(0, nb)

    def assembled(self, arrs, part):
        out = np.zeros_like(self.K)
        for q in range(self.R):
            for r in self.own_rows(q):
                out[r] = arrs[q][r]
        return np.tril(out) if part == "lower" else np.triu(out)

    # ------------------------------------------------------------------ solves
    def solve(self, rhs):
        R = self.R
        y = [np.zeros(self.Np) for _ in range(R)]
        for k in range(self.nblk):                                   # forward: L y = r, fan-out over row blocks
            c0, nb, o = self.blk(k)
            Fo = self.F[o]
            t = rhs[c0:c0 + nb] - Fo[c0:c0 + nb, :c0] @ y[o][:c0]
            yk = np.linalg.solve(np.tril(Fo[c0:c0 + nb, c0:c0 + nb]), t)
            for q in range(R):
                y[q][c0:c0 + nb] = yk                                 # broadcast
        a = [np.zeros(self.Np) for _ in range(R)]
        for k in reversed(range(self.nblk)):                         # backward: U a = y on the U = L' rows
            c0, nb, o = self.blk(k)
            Fo = self.F[o]
            t = y[o][c0:c0 + nb] - Fo[c0:c0 + nb, c0 + nb:] @ a[o][c0 + nb:]
            ak = np.linalg.solve(np.tril(Fo[c0:c0 + nb, c0:c0 + nb]).T, t)
            for q in range(R):
                a[q][c0:c0 + nb] = ak
        return a[0]

    # ------------------------------------------------------------------ inverse: backward sweep over row panels
    def inverse(self):
        R, Np = self.R, self.Np
        for q in range(R):
            for r in self.own_rows(q):
                self.G[q][r, :] = 0.0
        # W_kk of the own diagonal blocks: strictly-lower part -> G lower, its transpose -> G upper; diagonal TILES stay in
        # Dinv / DinvT (the GEMM substitutes them)
        for k in range(self.nblk):
            c0, nb, o = self.blk(k)
            W = np.linalg.inv(np.tril(self.F[o][c0:c0 + nb, c0:c0 + nb]))
            for bi in range(nb // T):
                for bj in range(nb // T):
                    blk = W[bi * T:(bi + 1) * T, bj * T:(bj + 1) * T]
                    if bi > bj:
                        self.G[o][c0 + bi * T:c0 + (bi + 1) * T, c0 + bj * T:c0 + (bj + 1) * T] = blk
                        self.G[o][c0 + bj * T:c0 + (bj + 1) * T, c0 + bi * T:c0 + (bi + 1) * T] = blk.T
        for k in reversed(range(self.nblk)):
            j0, nb, o = self.blk(k)
            r1 = j0 + nb
            Go = self.G[o]
            # 1. owner finalises X_J: columns beyond the block  Xt[c, i] = -sum_k accT[c, k] Wt_JJ[i, k]
            Wt = self.wt_block(o, j0, nb)                             # rows of Wt_JJ (upper triangular)
            P0 = [np.full((Np, self.NB), np.nan) for _ in range(R)]
            if r1 < Np:
                accT = Go[j0:r1, r1:].T.copy()                        # transpose kernel
                Xt = -nt(accT, Wt)                                    # NT GEMM, GEMM_KLO_N (k >= tile of i)
                Go[j0:r1, r1:] = Xt.T                                 # transposed-copy epilogue: X_J rows in place
                P0[o][r1:, :nb] = Xt
            P0[o][j0:r1, :nb] = Wt.T                                  # pack_wblock: Xt[c in J, i] = W_JJ[c, i] (lower)
            for q in range(R):                                        # 2. broadcast of the column-panel form
                if q != o:
                    P0[q][j0:, :nb] = P0[o][j0:, :nb]
            for q in range(R):
                XR = np.full((self.NB, Np), np.nan)
                XR[:nb, j0:] = P0[q][j0:, :nb].T                      # 3. local transpose: row-panel form X_J
                # (1) TRTRI update: own rows i < j0:  G[i, c >= j0] += U[i, J] X_J  == NT(F[i, Jcols], P0[c, :])
                rows = self.own_rows(q, 0, j0)
                if rows:
                    self.G[q][np.ix_(rows, range(j0, Np))] += nt(self.F[q][np.ix_(rows, range(j0, r1))], P0[q][j0:, :nb])
                # (2) LAUUM: own rows i >= j0: Kinv[i, J] = sum_{c >= i} Wt[i, c] X_J[j, c]   (lower part only)
                for i in self.own_rows(q, j0):
                    wrow = self.wt_row(q, i)                           # Wt row i: zeros left of the diagonal tile
                    jhi = min(i, r1 - 1)
                    lo = (i // T) * T
                    val = nt(wrow[None, lo:], XR[:jhi - j0 + 1, lo:])[0]
                    self.G[q][i, j0:jhi + 1] = val

    def wt_block(self, o, j0, nb):
        Wt = np.zeros((nb, nb))
        for bi in range(nb // T):
            for bj in range(nb // T):
                if bj > bi:
                    Wt[bi * T:(bi + 1) * T, bj * T:(bj + 1) * T] = self.G[o][j0 + bi * T:j0 + (bi + 1) * T, j0 + bj * T:j0 + (bj + 1) * T]
                elif bj == bi:
                    Wt[bi * T:(bi + 1) * T, bj * T:(bj + 1) * T] = self.DinvT[o][j0 + bi * T:j0 + (bi + 1) * T]
        return Wt

    def wt_row(self, q, i):
        row = np.zeros(self.Np)
        t0 = (i // T) * T
        row[t0 + T:] = self.G[q][i, t0 + T:]
        row[t0:t0 + T] = self.DinvT[q][i]
        return row

    # ------------------------------------------------------------------ predict: Vt = Kst L^-T, own columns maintained
    def predict_var(self, Kst, kdiag):
        R = self.R
        M = Kst.shape[0]
        Vt = [Kst.copy() for _ in range(R)]
        vacc = [np.zeros(M) for _ in range(R)]
        for k in range(self.nblk):
            c0, nb, o = self.blk(k)
            r1 = c0 + nb
            Lkk = np.tril(self.F[o][c0:r1, c0:r1])
            Pv = np.linalg.solve(Lkk, Vt[o][:, c0:r1].T).T             # owner: block TRSM (in place on its copy)
            for q in range(R):                                         # broadcast Pv
                vacc[q] += np.sum(Pv * Pv, axis=1)
                rows = self.own_rows(q, r1)
                if rows:                                               # own later columns: Vt[:, r] -= Pv L[r, c0:r1]'
                    Vt[q][:, rows] -= nt(Pv, self.F[q][np.ix_(rows, range(c0, r1))])
        return kdiag - vacc[0], vacc


def _spd(n, seed):
    rng = np.random.default_rng(seed)
    X = rng.standard_normal((n, 3))
    d2 = ((X[:, None, :] - X[None, :, :]) ** 2).sum(-1)
    return np.exp(-0.5 * d2) + 0.3 * np.eye(n), X


@pytest.mark.parametrize("R,rb,n", [(1, 2, 24), (2, 1, 20), (3, 2, 44), (4, 2, 36), (2, 4, 40), (8, 1, 36)])
def test_sharded_schedules_model(R, rb, n):
    K, X = _spd(n, 7 * R + rb)
    m = Model(K, R, rb)
    m.cholesky()
    L = np.linalg.cholesky(K)
    assert np.allclose(m.assembled(m.F, "lower"), L, atol=1e-12)
    assert np.allclose(np.triu(m.assembled(m.F, "upper"), 1)[:, :], np.triu(L.T, 1) * _beyond_block_mask(m), atol=1e-12)
    assert abs(np.sum(m.logd[0]) - np.linalg.slogdet(K)[1]) < 1e-10
    rhs = np.random.default_rng(1).standard_normal(n)
    assert np.allclose(m.solve(rhs), np.linalg.solve(K, rhs), atol=1e-10)
    m.inverse()
    assert np.allclose(m.assembled(m.G, "lower"), np.tril(np.linalg.inv(K)), atol=1e-10)
    Xs = np.random.default_rng(2).standard_normal((5, 3))
    Kst = np.exp(-0.5 * ((Xs[:, None, :] - X[None, :, :]) ** 2).sum(-1))
    var, vacc = m.predict_var(Kst, np.ones(5))
    V = np.linalg.solve(L, Kst.T)
    assert np.allclose(var, 1.0 - np.sum(V * V, axis=0), atol=1e-10)
    assert all(np.array_equal(vacc[0], v) for v in vacc)


def _beyond_block_mask(m):
    """U rows are kept only beyond the row's own diagonal block (inside it the upper part is scratch)"""
    mask = np.zeros((m.Np, m.Np))
    for r in range(m.Np):
        k = r // m.NB
        mask[r, min((k + 1) * m.NB, m.Np):] = 1.0
    return mask
