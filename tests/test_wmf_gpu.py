"""GPU parity tests of the WMF ALS path (pytest -m gpu), through the C ABI.

Bar (BASELINE.json): factors within 1e-4 relative of the reference on the same init.  "Relative" is
max|got - want| / max|want| per matrix.  The reference solves each row exactly (LU); the kernels run conjugate
gradient to a relative residual of 1e-6 (f32) / 1e-10 (f64):
  * float64 arithmetic: <= 1e-7 against the compiled reference's recorded factors (tests/golden) after every epoch;
  * float32 arithmetic (the throughput mode): <= 1e-4 after 1, 2 and 3 epochs on the goldens, and after 2 epochs on
    the ml-1m-shaped config C2 (K=64) against the oracle.
"""
import ctypes as C

import numpy as np
import pytest
from scipy import sparse

from conftest import golden

pytestmark = pytest.mark.gpu


def _rel(got, want):
    return float(np.abs(got - want).max() / np.abs(want).max())


def _csr(g):
    U, I, K = (int(v) for v in g["shape"])
    return sparse.csr_matrix((np.ones(g["indices"].shape[0]), g["indices"], g["indptr"]), shape=(U, I)), U, I, K


@pytest.mark.parametrize("name", ["wmf_small", "wmf_k64"])
@pytest.mark.parametrize("dtype,tol,short", [("float64", 1e-7, None), ("float32", 1e-4, "0"), ("float32", 1e-4, "112"),
                                             ("float32", 1e-4, "dual"), ("float32", 1e-4, "cta")])
def test_fit_matches_reference_golden(name, dtype, tol, short, monkeypatch):
    """short = "0": every non-empty row goes through the one-pass tensor-core solver (cymf_als_rows_tc_dev);
    "112": rows of <= 112 entries -- all rows of these small fixtures -- take the streaming CG kernel;
    "dual" (the default): rows of <= 128 (K = 128) / 64 entries are solved in their dual form on tensor-core tiles
    (cymf_als_rows_dual_dev), the rest by the one-pass solver."""
    import cymf_b200 as cymf
    if short == "dual":
        monkeypatch.setenv("CYMF_ALS_DUAL", "1")
    elif short == "0":
        monkeypatch.setenv("CYMF_ALS_DUAL", "0")             # every row through the warp-specialised one-pass solver
        monkeypatch.setenv("CYMF_ALS_SHORT", "0")
    elif short == "cta":
        monkeypatch.setenv("CYMF_ALS_DUAL", "0")             # ... and through the CTA-per-row form of it (als_tc.cu)
        monkeypatch.setenv("CYMF_ALS_SHORT", "0")
        monkeypatch.setenv("CYMF_ALS_WS", "0")
    elif short is not None:
        monkeypatch.setenv("CYMF_ALS_DUAL", "0")
        monkeypatch.setenv("CYMF_ALS_SHORT", short)
    g = golden(name + ".npz")
    X, U, I, K = _csr(g)
    m = cymf.WMF(K, float(g["wd"]), float(g["weight"]), dtype=dtype)
    m.W, m.H = g["W0"].copy(), g["H0"].copy()
    m.fit(X, 1, 4, verbose=False)
    print(name, dtype, "epoch 1 rel err", _rel(m.W, g["W_e1"]), _rel(m.H, g["H_e1"]), "cg iters", m.cg_iterations_)
    assert _rel(m.W, g["W_e1"]) <= tol and _rel(m.H, g["H_e1"]) <= tol
    m.fit(X, int(g["epochs"]) - 1, 4, verbose=False)          # warm start continues from W, H (wmf.pyx:88-92)
    print(name, dtype, "final rel err", _rel(m.W, g["W"]), _rel(m.H, g["H"]))
    assert _rel(m.W, g["W"]) <= tol and _rel(m.H, g["H"]) <= tol
    assert not m.W[3].any() and not m.H[5].any()              # rows without interactions are zeroed (wmf.pyx:154-156)
    assert m.cg_unconverged_ == 0


def test_als_typed_boundary_host_buffers(oracle):
    """WMF._als(indptr, indices, X, Y, num_threads) with host arrays, one half sweep, vs the oracle."""
    import cymf_b200 as cymf
    g = golden("wmf_k64.npz")
    X, U, I, K = _csr(g)
    for dtype, tol in (("float64", 1e-8), ("float32", 2e-5)):
        W, H = g["W0"].copy(), g["H0"].copy()
        Wo = W.copy()
        oracle.als_half(X.indptr, X.indices, Wo, H, float(g["wd"]), float(g["weight"]))
        m = cymf.WMF(K, float(g["wd"]), float(g["weight"]), dtype=dtype)
        m._als(X.indptr, X.indices, W, H, 8)
        assert _rel(W, Wo) <= tol
        assert m.cg_iterations_ > 0


def test_gram_kernel(oracle):
    import torch
    from cymf_b200 import _lib
    rng = np.random.default_rng(0)
    for n, K in ((1000, 20), (5000, 64), (777, 128), (3, 8)):
        Y = rng.normal(size=(n, K))
        want = Y.T @ Y + 0.01 * np.eye(K)
        for dtype, tol in ((_lib.F64, 1e-13), (_lib.F32, 2e-6)):
            dY = _lib.upload_factor(Y, dtype, torch.device("cuda"))
            nws = int(_lib.lib().cymf_gram_workspace_doubles(n, K))
            ws = torch.empty(nws, dtype=torch.float64, device="cuda")
            g64 = torch.empty(K * K, dtype=torch.float64, device="cuda")
            _lib.check(_lib.lib().cymf_gram_dev(_lib.ptr(dY), dtype, n, K, dY.shape[1], 0.01, 1, _lib.ptr(ws), nws,
                                                _lib.ptr(g64), None, None))
            got = g64.cpu().numpy().reshape(K, K)
            assert _rel(got, want) <= tol
            assert np.array_equal(got, got.T)                  # bitwise symmetric (fixed summation order)


@pytest.mark.parametrize("K", [32, 64, 96, 128])
def test_dual_row_solver_against_float64(K):
    """cymf_als_rows_dual_dev alone: rows of 0..128 entries (0..64 when ld < 128) in the three tile classes, against
    a float64 solve of the reference's row system in the transformed coordinates,
    (I + (w-1) sum y~ y~^T) x~ = w sum y~ (cymf/wmf.pyx:161-168 with G = I)."""
    import torch
    from cymf_b200 import _lib
    rng = np.random.default_rng(K)
    n_items, w = 5000, 10.0
    Y = rng.normal(size=(n_items, K)) * (0.6 / np.sqrt(K))               # |y~|^2 ~ 0.36: well inside Y~ Y~^T <= I
    top = 128 if K == 128 else 64
    lens = np.concatenate([rng.integers(65, top + 1, 70) if top == 128 else np.empty(0, np.int64), [top, top - 1],
                           rng.integers(33, 65, 90), [64, 33], rng.integers(0, 33, 150), [0, 1, 32, 0, 31]]).astype(np.int64)
    lens = np.sort(lens)[::-1].copy()                                     # heaviest first, as AlsSession deals them
    n128, n64, n32 = int((lens > 64).sum()), int(((lens > 32) & (lens <= 64)).sum()), int((lens <= 32).sum())
    indptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    indices = np.concatenate([np.sort(rng.choice(n_items, n, replace=False)) for n in lens] + [np.empty(0, np.int64)]).astype(np.int32)
    rows = lens.shape[0]
    want = np.zeros((rows, K))
    for r in range(rows):
        Yr = Y[indices[indptr[r]:indptr[r + 1]]]
        if Yr.shape[0]:
            want[r] = np.linalg.solve(np.eye(K) + (w - 1) * Yr.T @ Yr, w * Yr.sum(0))
    dY = torch.from_numpy(Y).to("cuda", torch.float32).contiguous()
    dX = torch.full((rows, K), 7.0, dtype=torch.float32, device="cuda")   # stale content must be overwritten
    d_ip, d_ix = torch.from_numpy(indptr).cuda(), torch.from_numpy(indices).cuda()
    order = torch.arange(rows, dtype=torch.int32, device="cuda")
    queue = torch.zeros(4, dtype=torch.int32, device="cuda")
    stats = torch.zeros(2, dtype=torch.int64, device="cuda")
    for rep in range(2):                                                  # twice: the work queues are reset per call
        _lib.check(_lib.lib().cymf_als_rows_dual_dev(_lib.ptr(d_ip), _lib.ptr(d_ix), _lib.ptr(order), n128, n64, n32,
                                                     _lib.ptr(dX), _lib.ptr(dY), _lib.F32, K, K, w, 1e-6, 256,
                                                     _lib.ptr(queue), _lib.ptr(stats), None))
        torch.cuda.synchronize()
        got = dX.double().cpu().numpy()
        Yf = dY.double().cpu().numpy()                                    # the f32 inputs the kernel actually saw
        err = np.abs(got - want).max() / np.abs(want).max()
        print(f"K={K} dual solver: rel err {err:.2e}, CG iterations/row {int(stats[0]) / rows / (rep + 1):.1f}")
        assert err <= 2e-5
        assert not got[lens == 0].any()
        assert int(stats[1]) == 0
    del Yf


@pytest.mark.parametrize("gather", ["tma", "cp.async"])
@pytest.mark.parametrize("K", [32, 64, 96, 128])
def test_ws_row_solver_against_float64(K, gather, monkeypatch):
    """cymf_als_rows_ws_dev alone (warp-specialised persistent solver + its host-side schedule): rows of 0 .. 1600
    entries -- partial chunks, exactly one chain of 512, several chains -- from a random warm start, against a
    float64 solve of (I + (w-1) sum y~ y~^T) x~ = w sum y~ (cymf/wmf.pyx:161-168 with G = I).  Both gathers: through
    the TMA unit (tile::gather4 over a tensor map of Y) and by cp.async."""
    import torch
    from cymf_b200 import _lib
    monkeypatch.setenv("CYMF_ALS_WS_TMA", "1" if gather == "tma" else "0")
    L = _lib.lib()
    rng = np.random.default_rng(100 + K)
    n_items, w = 6000, 10.0
    Y = rng.normal(size=(n_items, K))
    Y = np.linalg.solve(np.linalg.cholesky(Y.T @ Y + 0.01 * np.eye(K)), Y.T).T          # y~ = L^-1 y: Y~^T Y~ <= I
    lens = np.concatenate([[0, 1, 5, 31, 32, 33, 64, 100, 129, 200, 511, 512, 513, 700, 1024, 1025, 1600, 0],
                           rng.integers(1, 400, 700)]).astype(np.int64)
    rng.shuffle(lens)
    indptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    indices = np.concatenate([np.sort(rng.choice(n_items, n, replace=False)) for n in lens]).astype(np.int32)
    rows = lens.shape[0]
    want = np.zeros((rows, K))
    for r in range(rows):
        Yr = Y[indices[indptr[r]:indptr[r + 1]]]
        if Yr.shape[0]:
            want[r] = np.linalg.solve(np.eye(K) + (w - 1) * Yr.T @ Yr, w * Yr.sum(0))
    n_ctas = int(L.cymf_als_ws_ctas())
    assert n_ctas > 0
    row_ids = np.arange(rows, dtype=np.int32)
    cta_ptr, rowinfo = np.empty(n_ctas + 1, np.int32), np.empty(4 * rows, np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    _lib.check(L.cymf_als_ws_schedule_host(p(indptr), p(row_ids), rows, n_ctas, 256, p(cta_ptr), p(rowinfo)))
    assert cta_ptr[0] == 0 and cta_ptr[-1] == rows and np.array_equal(np.sort(rowinfo[0::4]), row_ids)
    load = np.add.reduceat(np.concatenate([rowinfo[1::4] + 256, [0]]), np.minimum(cta_ptr[:-1], rows))[np.diff(cta_ptr) > 0]
    assert load.max() <= load.min() + 1600 + 256                                        # LPT: bins differ by one row at most
    dY = torch.from_numpy(Y).to("cuda", torch.float32).contiguous()
    X0 = rng.normal(size=(rows, K)) * 0.05
    dX = torch.from_numpy(X0).to("cuda", torch.float32).contiguous()
    d_ix = torch.from_numpy(indices).cuda()
    d_ri, d_cp = torch.from_numpy(rowinfo).cuda(), torch.from_numpy(cta_ptr).cuda()
    stats = torch.zeros(2, dtype=torch.int64, device="cuda")
    debug = torch.zeros(32, dtype=torch.int64, device="cuda")
    for rep in range(2):                                                                # second pass: warm start = the solution
        _lib.check(L.cymf_als_rows_ws_dev(_lib.ptr(d_ri), _lib.ptr(d_cp), n_ctas, _lib.ptr(d_ix), _lib.ptr(dX), _lib.ptr(dY),
                                          n_items, _lib.F32, K, K, w, 1e-6, 256, _lib.ptr(stats),
                                          _lib.ptr(debug), None))
        torch.cuda.synchronize()
        assert int(debug[0]) == 0, f"hand-over timed out: {debug.tolist()}"
        got = dX.double().cpu().numpy()
        err = np.abs(got - want).max() / np.abs(want).max()
        print(f"K={K} ws solver pass {rep}: rel err {err:.2e}, CG iterations so far {int(stats[0])}")
        assert err <= 2e-5
        assert not got[lens == 0].any()
        assert int(stats[1]) == 0


def test_c2_shape_f32_within_1e4(oracle):
    """Config C2: K=64 on the ml-1m-shaped matrix, 2 epochs, float32 CG vs the oracle's exact solves."""
    import cymf_b200 as cymf
    train, _ = cymf.synth.movielens_like("ml-1m")
    Wo, Ho = oracle.wmf_fit(train, 64, 0.01, 10.0, 2)
    m = cymf.WMF(64, 0.01, 10.0)
    m.fit(train, 2, 8, verbose=False)
    rows = train.shape[0] + train.shape[1]
    print("C2 rel err", _rel(m.W, Wo), _rel(m.H, Ho), "mean CG iterations/row", m.cg_iterations_ / (2 * rows))
    assert _rel(m.W, Wo) <= 1e-4 and _rel(m.H, Ho) <= 1e-4
    assert m.cg_unconverged_ == 0


def test_evaluator_hook_and_early_stopping():
    import cymf_b200 as cymf
    train, test = cymf.synth.movielens_like("ml-100k")
    ev = cymf.evaluator.AverageOverAllEvaluator(test, train, k=5)
    m = cymf.WMF(20, 0.01, 10.0)
    m.fit(train, 4, 8, valid_evaluator=ev, early_stopping=True, verbose=False)
    assert m.valid_dcg > 0.2 and m.W.shape == (943, 20) and m.H.shape == (1682, 20)
    with pytest.raises(ValueError):
        cymf.WMF().fit(train, early_stopping=True)
    with pytest.raises(ValueError):
        cymf.WMF().fit(None)


@pytest.mark.parametrize("K", [32, 64, 128])
def test_heavy_rows_direct_solve(oracle, K, monkeypatch):
    """Rows of >= heavy_min entries take the direct path (gathered tcgen05 Gram per 512-entry slab + f64 LDL^T solve,
    cymf_als_heavy_rows_dev) instead of CG: same factors as the reference's dgesv solves, and as the CG path.
    (The tail divisor is set explicitly: by default the warp-specialised solver keeps every row that does not exceed
    its CTA's fair share by 40 k entries, which no row of this small matrix does.)"""
    import cymf_b200 as cymf
    monkeypatch.setenv("CYMF_ALS_TAIL_DIVISOR", "256")
    X = cymf.synth.synth_implicit(900, 700, 60000, seed=12).tolil()
    X[3, :] = 1                                               # a row with every item: 700 entries = 2 slabs
    X[:, 5] = 1                                               # and a column with every user: 900 entries
    X = X.tocsr()
    Wo, Ho = oracle.wmf_fit(X, K, 0.01, 10.0, 2)
    out = {}
    for heavy_min in (0, 100):
        m = cymf.WMF(K, 0.01, 10.0, heavy_min=heavy_min)
        m.fit(X, 2, 1, verbose=False)
        out[heavy_min] = (m.W.copy(), m.H.copy())
        for got, want in ((m.W, Wo), (m.H, Ho)):
            assert np.abs(got - want).max() <= 1e-4 * np.abs(want).max(), (heavy_min, np.abs(got - want).max())
    assert np.abs(out[0][0] - out[100][0]).max() <= 2e-5 * np.abs(Wo).max()


def test_device_generated_matrix_normal_equations(monkeypatch):
    """The C5 pipeline at 1/100 scale (100 k x 10 k, ~8.6 M nnz; tools/c5_als.py): matrix generated on the device,
    X^T / deal / blocks by prep.cu, factors initialised on the device, and after each half sweep the sampled rows
    (plus the heaviest one) satisfy the reference's normal equations (cymf/wmf.pyx:161-168) in f64."""
    import os
    import sys
    import torch
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import c5_als
    from cymf_b200.synth import synth_implicit_device
    from cymf_b200.wmf import AlsSession
    U, I, K = 100_000, 10_000, 64
    ip, ix = synth_implicit_device(U, I, 10_000_000, seed=104)
    g = torch.Generator(device="cuda")
    g.manual_seed(4321)
    W = (torch.rand((U, K), device="cuda", generator=g) * 0.2 - 0.1) / K
    H = (torch.rand((I, K), device="cuda", generator=g) * 0.2 - 0.1) / K
    monkeypatch.setenv("CYMF_ALS_TAIL_DIVISOR", "4096")    # rows of >= 2.1 k entries take the direct solve here
    sess = AlsSession((ip, ix, (U, I)), W, H, 0.01, 10.0, K=K, heavy_min=2048)
    assert sess.heavy_i is not None and sess.heavy_i[0] > 10
    sess.epoch()
    sess.user_half()
    assert c5_als.residuals(sess, "user", 12, seed=0) <= 1e-4
    sess.item_half()
    assert c5_als.residuals(sess, "item", 12, seed=1) <= 1e-4
    assert sess.stats()[1] == 0                            # no row stopped at cg_max_iter


@pytest.mark.parametrize("dtype,ld,tol", [("float32", 128, 2e-6), ("float32", 64, 2e-6), ("float64", 20, 1e-13)])
def test_rows_times_matrix_in_place_multi_tile(dtype, ld, tol):
    """cymf_rows_times_matrix_dev with out == in (how the warm start and the back-transform are applied) over several
    128-row tiles: same result as out-of-place, and both match NumPy.  (tcgen05 path for f32 with ld % 32 == 0, FFMA
    path otherwise.)"""
    import torch
    from cymf_b200 import _lib
    rng = np.random.default_rng(ld)
    rows = 1000
    tdt = torch.float32 if dtype == "float32" else torch.float64
    X = rng.normal(size=(rows, ld))
    B = rng.normal(size=(ld, ld)) / np.sqrt(ld)
    dX = torch.from_numpy(X).to("cuda", tdt)
    dB = torch.from_numpy(B).to("cuda", tdt).contiguous()
    out = torch.empty_like(dX)
    code = _lib.DTYPES[dtype]
    _lib.check(_lib.lib().cymf_rows_times_matrix_dev(_lib.ptr(dX), _lib.ptr(out), _lib.ptr(dB), code, rows, ld, None))
    inplace = dX.clone()
    _lib.check(_lib.lib().cymf_rows_times_matrix_dev(_lib.ptr(inplace), _lib.ptr(inplace), _lib.ptr(dB), code, rows, ld, None))
    torch.cuda.synchronize()
    assert torch.equal(out, inplace)
    want = dX.double().cpu().numpy() @ dB.double().cpu().numpy()
    assert _rel(out.double().cpu().numpy(), want) <= tol
