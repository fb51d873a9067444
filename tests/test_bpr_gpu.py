"""GPU parity tests of the BPR path (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI.

Tolerances (written here on purpose):
  * serialized f64 replay vs the compiled reference's own outputs (tests/golden): BIT-EXACT.  The kernel
    performs the reference's operations in the reference's order with IEEE-rounded *, +, -, /, sqrt and
    evaluates exp() with the C library's own table-driven algorithm (csrc/bpr.cu exp_libm), so every
    element of W and H must compare equal.
  * Hogwild kernel run one triplet at a time (max_inflight=1, f64, plain stores) vs the oracle fed the same
    Philox negatives: same arithmetic except the dot product is tree-reduced -> 1e-12 relative.
  * Hogwild f32 at full concurrency: ranking metrics within 1 % (relative) of the oracle's, BASELINE.json.
"""
import ctypes as C

import numpy as np
import pytest
from scipy import sparse

from conftest import golden

pytestmark = pytest.mark.gpu


def _csr(g):
    U, I, K = (int(v) for v in g["shape"])
    X = sparse.csr_matrix((np.ones(g["indices"].shape[0]), g["indices"], g["indptr"]), shape=(U, I))
    return X, U, I, K


def _ulp_report(got, want):
    same = float((got == want).mean())
    err = float(np.abs(got - want).max())
    return same, err


@pytest.mark.parametrize("name", ["bpr_sgd", "bpr_adagrad", "bpr_adam", "bpr_sgd_mid"])
def test_replay_matches_reference_golden(name):
    import cymf_b200 as cymf
    g = golden(name + ".npz")
    X, U, I, K = _csr(g)
    m = cymf.BPR(K, float(g["lr"]), str(g["opt"]), float(g["wd"]), mode="replay")
    m.W, m.H = g["W0"].copy(), g["H0"].copy()
    m.valid_evaluator, m.early_stopping = None, False
    m._fit_bpr(g["users"], g["positives"], X, int(g["epochs"]), float(g["lr"]), float(g["wd"]), 1, False)
    for got, want in ((m.W, g["W"]), (m.H, g["H"])):
        same, err = _ulp_report(got, want)
        print(f"{name}: bit-identical fraction {same:.6f}, max|diff| {err:.3e}")
        assert np.array_equal(got, want)


def test_replay_full_fit_equals_oracle_c1_shape(oracle):
    """Config C1 (943 x 1682, K=20, lr=0.01, wd=0.01) through the public fit(): same prologue, same stream."""
    import cymf_b200 as cymf
    X = cymf.synth.synth_implicit(943, 1682, 100_000, seed=100)
    train, _ = cymf.synth.split_train_test(X, 100)
    epochs = 3
    m = cymf.BPR(20, 0.01, "adam", 0.01, mode="replay")
    m.fit(train, num_epochs=epochs, num_threads=1, verbose=False)
    Xc, W, H, users, positives = oracle.bpr_prologue(train, 20)
    rec = oracle.bpr_fit(W, H, users, positives, Xc, epochs, 0.01, 0.01, "adam", record=True)
    assert m.n_applied_ == int(rec["applied"].sum())
    assert m.n_attempted_ == epochs * train.nnz
    for got, want in ((m.W, W), (m.H, H)):
        same, err = _ulp_report(got, want)
        print(f"C1 adam replay: bit-identical fraction {same:.6f}, max|diff| {err:.3e}")
        assert np.array_equal(got, want)


@pytest.mark.parametrize("mode", [2])
def test_host_abi_replay(oracle, mode):
    """cymf_bpr_fit_host: HOST buffers in, HOST buffers out (what a cgo/JNI/ctypes binding would call)."""
    from cymf_b200 import _lib
    g = golden("bpr_adagrad.npz")
    X, U, I, K = _csr(g)
    W, H = g["W0"].copy(), g["H0"].copy()
    applied = C.c_int64(0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    users, pos = np.ascontiguousarray(g["users"]), np.ascontiguousarray(g["positives"])
    ip, ix = np.ascontiguousarray(X.indptr, np.int32), np.ascontiguousarray(X.indices, np.int32)
    _lib.check(_lib.lib().cymf_bpr_fit_host(p(W), p(H), U, I, K, p(users), p(pos), users.shape[0], p(ip), p(ix),
                                            int(g["epochs"]), float(g["lr"]), float(g["wd"]), _lib.ADAGRAD, mode,
                                            1234, C.byref(applied)))
    assert np.array_equal(W, g["W"]) and np.array_equal(H, g["H"])
    assert 0 < applied.value <= users.shape[0] * int(g["epochs"])


@pytest.mark.parametrize("opt", ["sgd", "adagrad", "adam"])
@pytest.mark.parametrize("K", [20, 64, 128, 300])
def test_hogwild_kernel_one_triplet_at_a_time(oracle, opt, K):
    """The throughput kernel (every lane-group shape) is arithmetically the reference update."""
    import cymf_b200 as cymf
    from cymf_b200 import _lib
    U, I, nnz, epochs, seed = 80, 120, 1500, 2, 77
    X = cymf.synth.synth_implicit(U, I, nnz, seed=3)
    m = cymf.BPR(K, 0.05, opt, 0.01, dtype="float64", scatter="store", seed=seed, max_inflight=1)
    m.fit(X, num_epochs=epochs, verbose=False)
    Xc, W, H, users, positives = oracle.bpr_prologue(X, K)
    N = users.shape[0]
    neg = np.empty(epochs * N, np.int32)
    for e in range(epochs):
        _lib.check(_lib.lib().cymf_bpr_negatives_host(seed, e, 0, N, I, neg[e * N:].ctypes.data_as(C.c_void_p)))
    assert neg.min() >= 0 and neg.max() < I and len(np.unique(neg)) > I // 2
    rec = oracle.bpr_fit(W, H, users, positives, Xc, epochs, 0.05, 0.01, opt, negatives=neg, record=True)
    assert m.n_applied_ == int(rec["applied"].sum())
    assert np.abs(m.W - W).max() <= 1e-12 * np.abs(W).max()
    assert np.abs(m.H - H).max() <= 1e-12 * np.abs(H).max()


@pytest.mark.parametrize("opt", ["sgd", "adagrad", "adam"])
def test_hogwild_red_scatter_equals_store_when_serial(oracle, opt):
    """scatter='red' (parameter steps and optimizer-state increments applied with red.global.add) is the same update
    as plain stores when one triplet is in flight."""
    import cymf_b200 as cymf
    X = cymf.synth.synth_implicit(80, 120, 1500, seed=3)
    out = []
    for scatter in ("store", "red"):
        m = cymf.BPR(64, 0.05, opt, 0.01, dtype="float64", scatter=scatter, max_inflight=1)
        m.fit(X, num_epochs=2, verbose=False)
        out.append((m.W.copy(), m.H.copy()))
    assert np.abs(out[0][0] - out[1][0]).max() <= 1e-11
    assert np.abs(out[0][1] - out[1][1]).max() <= 1e-11


@pytest.mark.parametrize("opt,lr,floor", [("adagrad", 0.05, 0.9), ("adam", 0.01, 0.5)])
def test_hogwild_state_survives_heavy_collisions(oracle, opt, lr, floor):
    """943 user rows with ~40 triplets of each in flight (max_inflight=0 fills the machine; the default cap is 1024
    triplets): with reductions on the optimizer state and Adam's serial-bound clamp the factors stay finite.  AdaGrad
    ranks as well as the reference; Adam, whose first moment is advanced by 40 stale increments at once, keeps
    70-80 % of the reference's metrics (measured over several runs; the races make it vary) where plain stores
    diverge to inf -- the floors below only separate "degraded by staleness" from "diverged"."""
    import cymf_b200 as cymf
    train, test = cymf.synth.movielens_like("ml-100k")
    Xc, W, H, users, positives = oracle.bpr_prologue(train, 20)
    oracle.bpr_fit(W, H, users, positives, Xc, 20, lr, 0.01, opt)
    want = _mean_metrics(oracle, W, H, test, train)
    m = cymf.BPR(20, lr, opt, 0.01, max_inflight=0)
    m.fit(train, num_epochs=20, verbose=False)
    assert np.isfinite(m.W).all() and np.isfinite(m.H).all()
    got = _mean_metrics(oracle, m.W, m.H, test, train)
    print(opt, "reference", want, "gpu, machine-filling concurrency", got)
    for k in want:
        assert got[k] >= floor * want[k], (k, got[k], want[k])


def test_validation_on_device_equals_host_path():
    """fit(valid_evaluator=..., early_stopping=True): cymf_b200's evaluator scores the device-resident factors and the
    best epoch is snapshotted on the device; any other evaluator object gets NumPy arrays as in the reference
    (bpr.pyx:173-190).  Both paths must select the same epoch and return the same factors (replay mode is
    deterministic)."""
    import cymf_b200 as cymf
    train, test = cymf.synth.movielens_like("ml-100k")
    ev = cymf.evaluator.AverageOverAllEvaluator(test, train, k=5)

    class Foreign:                                  # not an instance of cymf_b200.Evaluator -> host path
        def __init__(self, inner):
            self.inner, self.calls = inner, 0

        def evaluate(self, W, H):
            assert isinstance(W, np.ndarray) and isinstance(H, np.ndarray)
            self.calls += 1
            return self.inner.evaluate(W, H)

    sub = train[:200]
    sub_test = test[:200]
    ev_small = cymf.evaluator.AverageOverAllEvaluator(sub_test, sub, k=5)
    out = []
    for evaluator in (ev_small, Foreign(ev_small)):
        m = cymf.BPR(8, 0.05, "sgd", 0.01, mode="replay")
        m.fit(sub, num_epochs=4, valid_evaluator=evaluator, early_stopping=True, verbose=False)
        out.append((m.W.copy(), m.H.copy(), m.valid_dcg))
    assert out[0][2] == out[1][2] > 0
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    h = cymf.BPR(20, 0.01, "adam", 0.01)
    h.fit(train, num_epochs=5, valid_evaluator=ev, verbose=False)          # Hogwild + on-device validation
    assert h.valid_dcg > 0.05 and np.isfinite(h.W).all()


def _mean_metrics(oracle, W, H, test, train):
    rs = [oracle.evaluate(W, H, test, train, k=5, seed=s) for s in range(5)]     # optuna_example.py:63-65
    return {k: float(np.mean([r[k] for r in rs])) for k in rs[0]}


@pytest.mark.parametrize("opt,lr", [("sgd", 0.05), ("adam", 0.01)])
def test_hogwild_f32_metric_parity_c1(oracle, opt, lr):
    """Config C1 shape: Recall@5 / DCG@5 / MAP@5 of the concurrent f32 kernel within 1 % of the reference's."""
    # The reference's own metrics move by ~0.4 % (1 sigma) with the negative-sampling seed alone (measured
    # with the oracle: seeds 1234,1..5 span 1.0-1.9 %), so both sides are averaged over 6 seeds; the standard
    # error of the difference of means is then ~0.25 % and the 1 % bound is a 4-sigma test.
    import cymf_b200 as cymf
    train, test = cymf.synth.movielens_like("ml-100k")
    epochs, seeds = 30, (1234, 1, 2, 3, 4, 5)
    wants, gots = [], []
    for seed in seeds:
        Xc, W, H, users, positives = oracle.bpr_prologue(train, 20)
        oracle.bpr_fit(W, H, users, positives, Xc, epochs, lr, 0.01, opt, seed=seed)
        wants.append(_mean_metrics(oracle, W, H, test, train))
        m = cymf.BPR(20, lr, opt, 0.01, seed=seed)
        m.fit(train, num_epochs=epochs, num_threads=8, verbose=False)
        gots.append(_mean_metrics(oracle, m.W, m.H, test, train))
        assert 0.8 * m.n_attempted_ < m.n_applied_ < m.n_attempted_
    want = {k: float(np.mean([r[k] for r in wants])) for k in wants[0]}
    got = {k: float(np.mean([r[k] for r in gots])) for k in gots[0]}
    print(opt, "reference", want, "gpu", got)
    assert want["Recall@5"] > 0.3, "synthetic data too flat to detect regressions"
    for k in want:
        assert abs(got[k] - want[k]) <= 0.01 * want[k], (k, got[k], want[k])


@pytest.mark.parametrize("opt", ["sgd", "adam"])
def test_hogwild_metric_parity_c3_shape(opt):
    """The benchmarked scale (configs[2] shape: 138,493 x 26,744, ~18 M pairs, machine-filling concurrency -- no
    in-flight cap applies at this size): K=64, 3 epochs, concurrent f32 kernel vs the COMPILED reference under OpenMP
    on every host core; Recall@5 / DCG@5 / MAP@5 (5 evaluator seeds) within 1 % relative (tools/hogwild_parity_c3.py)."""
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if not os.path.isdir(os.path.join(root, "oracle", "_ref", "cymf")):
        pytest.skip("compiled reference (oracle/_ref) not built")
    sys.path.insert(0, os.path.join(root, "tools"))
    import hogwild_parity_c3
    res = hogwild_parity_c3.compare(64, opt, 3)
    print(res)
    assert res["reference"]["Recall@5"] > 0.1, "metrics too flat to detect a regression"
    assert res["max_rel_diff"] <= 0.01, res
