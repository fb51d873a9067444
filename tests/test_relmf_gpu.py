"""RelMF CUDA path vs the compiled reference's vectors and the CPU oracle (pytest -m gpu).

Bars: serialized replay bit-exact (np.array_equal) against factors recorded from the compiled reference
(tests/golden/relmf_*.npz) and against the oracle at a larger shape; the throughput kernel, one sample in flight,
within 1e-10 of the oracle fed the same Philox cells for every lane-group shape; the concurrent f32 kernel within
2 % of the oracle's ranking metrics (6-seed means)."""
import ctypes as C

import numpy as np
import pytest
from scipy import sparse

from conftest import golden

pytestmark = pytest.mark.gpu


def _csr(g):
    U, I, K = (int(v) for v in g["shape"])
    return sparse.csr_matrix((g["data"], g["indices"], g["indptr"]), shape=(U, I)), U, I, K


@pytest.mark.parametrize("name", ["relmf_sgd", "relmf_adagrad", "relmf_adam", "relmf_ratings"])
def test_replay_bitwise_vs_compiled_reference(name):
    import cymf_b200 as cymf
    g = golden(name + ".npz")
    X, U, I, K = _csr(g)
    m = cymf.RelMF(K, float(g["clip"]), float(g["lr"]), str(g["opt"]), float(g["wd"]), mode="replay")
    m.fit(X if name != "relmf_sgd" else X.toarray(), int(g["epochs"]), 1)          # dense input is accepted too
    assert np.array_equal(m.W, g["W"]) and np.array_equal(m.H, g["H"])
    assert m.n_samples_ == U * I * int(g["epochs"])


def test_replay_host_abi(oracle):
    """cymf_relmf_fit_host: HOST buffers in, HOST buffers out."""
    from cymf_b200 import _lib
    g = golden("relmf_ratings.npz")
    X, U, I, K = _csr(g)
    W, H = g["W0"].copy(), g["H0"].copy()
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    ip, ix = np.ascontiguousarray(X.indptr, np.int32), np.ascontiguousarray(X.indices, np.int32)
    val, prop = np.ascontiguousarray(X.data, np.float64), np.ascontiguousarray(g["propensities"], np.float64)
    _lib.check(_lib.lib().cymf_relmf_fit_host(p(W), p(H), U, I, K, p(ip), p(ix), p(val), p(prop), int(g["epochs"]),
                                              float(g["lr"]), float(g["wd"]), float(g["clip"]), _lib.ADAM, 2, 1234))
    assert np.array_equal(W, g["W"]) and np.array_equal(H, g["H"])
    # Hogwild modes run and move the factors in the same direction
    for mode in (0, 1):
        W1, H1 = g["W0"].copy(), g["H0"].copy()
        _lib.check(_lib.lib().cymf_relmf_fit_host(p(W1), p(H1), U, I, K, p(ip), p(ix), p(val), p(prop), 2,
                                                  float(g["lr"]), float(g["wd"]), float(g["clip"]), _lib.ADAM, mode, 7))
        assert np.isfinite(W1).all() and np.abs(W1 - g["W0"]).max() > 0


@pytest.mark.parametrize("opt", ["sgd", "adam"])
def test_replay_midsize_vs_oracle(oracle, opt):
    import cymf_b200 as cymf
    X = cymf.synth.synth_implicit(150, 220, 3000, seed=8)
    m = cymf.RelMF(20, 0.1, 0.01, opt, 0.01, mode="replay")
    m.fit(X, 2, 1)
    W, H = oracle.init_factors(150, 220, 20)
    oracle.relmf_fit(W, H, X, 2, 0.01, 0.01, 0.1, opt)
    assert np.array_equal(m.W, W) and np.array_equal(m.H, H)


@pytest.mark.parametrize("opt", ["sgd", "adagrad", "adam"])
@pytest.mark.parametrize("K", [20, 64, 128, 300])
def test_hogwild_kernel_one_sample_at_a_time(oracle, opt, K):
    """The throughput kernel (every lane-group shape) is arithmetically the reference update."""
    import cymf_b200 as cymf
    from cymf_b200 import _lib
    U, I, epochs, seed = 40, 60, 2, 77
    X = cymf.synth.synth_implicit(U, I, 500, seed=3).astype(np.float64)
    X.data[:] = np.random.default_rng(1).integers(1, 6, X.nnz) / 5.0                 # values matter
    m = cymf.RelMF(K, 0.2, 0.05, opt, 0.01, dtype="float64", scatter="store", seed=seed, max_inflight=1)
    m.fit(X, epochs, 1)
    n = U * I
    cells = np.empty(epochs * n, np.int64)
    for e in range(epochs):
        _lib.check(_lib.lib().cymf_relmf_cells_host(seed, e, 0, n, U, I, cells[e * n:].ctypes.data_as(C.c_void_p)))
    assert cells.min() >= 0 and cells.max() < n and len(np.unique(cells)) > n // 2
    W, H = oracle.init_factors(U, I, K)
    oracle.relmf_fit(W, H, X, epochs, 0.05, 0.01, 0.2, opt, cells=cells)
    # f64 with a tree-reduced dot instead of the sequential one: rounding-level differences only (Adam's
    # normalised steps of size lr = 0.05 let them grow to a few 1e-12 over 4,800 samples)
    assert np.abs(m.W - W).max() <= 1e-10 * np.abs(W).max()
    assert np.abs(m.H - H).max() <= 1e-10 * np.abs(H).max()


def test_philox_cells_are_uniform():
    """u and i drawn independently == r ~ U[0, U*I), u = r / I, i = r % I: chi-square on both marginals."""
    from cymf_b200 import _lib
    U, I, n = 37, 101, 400_000
    cells = np.empty(n, np.int64)
    _lib.check(_lib.lib().cymf_relmf_cells_host(5, 0, 0, n, U, I, cells.ctypes.data_as(C.c_void_p)))
    for counts, bins in ((np.bincount(cells // I, minlength=U), U), (np.bincount(cells % I, minlength=I), I)):
        chi2 = ((counts - n / bins) ** 2 / (n / bins)).sum()
        assert chi2 < bins + 6 * np.sqrt(2 * bins)


def test_hogwild_f32_metric_parity(oracle):
    """ml-100k shape: ranking metrics of the concurrent f32 kernel against the reference algorithm (oracle)."""
    import cymf_b200 as cymf
    train, test = cymf.synth.movielens_like("ml-100k")
    epochs, seeds = 6, (1234, 1, 2)
    wants, gots = [], []
    for seed in seeds:
        W, H = oracle.init_factors(train.shape[0], train.shape[1], 20)
        oracle.relmf_fit(W, H, train, epochs, 0.01, 0.001, 0.1, "adam", seed=seed)
        wants.append(np.mean([list(oracle.evaluate(W, H, test, train, k=5, seed=s).values()) for s in range(5)], 0))
        m = cymf.RelMF(20, 0.1, 0.01, "adam", 0.001, seed=seed)
        m.fit(train, epochs, 8)
        gots.append(np.mean([list(oracle.evaluate(m.W, m.H, test, train, k=5, seed=s).values()) for s in range(5)], 0))
    want, got = np.mean(wants, 0), np.mean(gots, 0)
    print("reference", want, "gpu", got)
    assert want.min() > 0.05, "synthetic data too flat to detect regressions"
    assert np.all(np.abs(got - want) <= 0.02 * want), (got, want)


def test_validation_and_early_stopping():
    import cymf_b200 as cymf
    train, test = cymf.synth.movielens_like("ml-100k")
    sub, sub_test = train[:150], test[:150]
    ev = cymf.evaluator.AverageOverAllEvaluator(sub_test, sub, k=5)
    m = cymf.RelMF(8, 0.1, 0.01, "adam", 0.001, mode="replay", samples_per_epoch=20000)
    m.fit(sub, 3, 1, valid_evaluator=ev, early_stopping=True)
    assert m.valid_dcg > 0 and np.isfinite(m.W).all()
