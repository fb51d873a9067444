"""CPU check of the algebra behind cymf_als_rows_dual_dev (cymf_b200/csrc/als_dual.cu): in the coordinates of the Cholesky
change of variables the reference's row system (cymf/wmf.pyx:161-168)

    (G + (w-1) sum_i y_i y_i^T) x = w sum_i y_i ,     G = Y^T Y + wd I = L L^T ,  y~ = L^-1 y ,  x~ = L^T x

is (I_K + c Y~^T Y~) x~ = w Y~^T 1 with c = w - 1, and by the push-through identity the same x~ is w Y~^T z with
(I_n + c Y~ Y~^T) z = 1: an n x n system whose eigenvalues lie in [1, w] because 0 <= Y~ Y~^T <= I.  Checked here in
float64 against the oracle's restatement of the reference loop on a small matrix."""
import numpy as np
from scipy import sparse


def test_dual_form_equals_the_reference_row_solve(oracle):
    rng = np.random.default_rng(7)
    U, I, K, w, wd = 40, 90, 24, 10.0, 0.01
    X = sparse.random(U, I, density=0.12, random_state=3, format="csr")
    X.data[:] = 1.0
    X = X.tolil(); X[5, :] = 0; X = X.tocsr(); X.eliminate_zeros()          # a row without entries
    W = rng.normal(size=(U, K)) * 0.1
    H = rng.normal(size=(I, K)) * 0.1
    want = W.copy()
    oracle.als_half(X.indptr, X.indices, want, H, wd, w)                    # the reference's loop (dgesv per row)
    G = H.T @ H + wd * np.eye(K)
    L = np.linalg.cholesky(G)
    Ht = np.linalg.solve(L, H.T).T                                          # y~ = L^-1 y
    got = np.zeros_like(W)
    worst_cond = 0.0
    for u in range(U):
        idx = X.indices[X.indptr[u]:X.indptr[u + 1]]
        n = idx.shape[0]
        if n == 0:
            continue                                                        # wmf.pyx:154-156: zeroed
        Yr = Ht[idx]                                                        # n x K
        M = np.eye(n) + (w - 1.0) * Yr @ Yr.T
        ev = np.linalg.eigvalsh(M)
        assert ev[0] >= 1.0 - 1e-12 and ev[-1] <= w + 1e-9                  # 0 <= Y~ Y~^T <= I
        worst_cond = max(worst_cond, ev[-1] / ev[0])
        z = np.linalg.solve(M, np.ones(n))
        xt = w * Yr.T @ z                                                   # x~
        primal = np.linalg.solve(np.eye(K) + (w - 1.0) * Yr.T @ Yr, w * Yr.sum(0))
        assert np.abs(xt - primal).max() <= 1e-11 * max(1.0, np.abs(primal).max())
        got[u] = np.linalg.solve(L.T, xt)                                   # x = L^-T x~
    assert worst_cond <= w
    assert np.abs(got - want).max() <= 1e-9 * np.abs(want).max()
    assert not got[5].any() and not want[5].any()
