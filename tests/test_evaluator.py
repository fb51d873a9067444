"""Evaluator parity: candidate lists (host side of the C ABI, runs without a GPU) and, on the GPU, ranks and
metrics against the oracle and against the compiled reference's recorded outputs (tests/golden/evaluator.npz).

Bars: candidate lists and ranked indices bit-exact; metric means within 1e-15 (the reference scores with a
BLAS `np.dot`, the oracle and the kernel with a k-ascending dot -- identical ranking, and per-user metric values
are ratios of small integers and log2 terms, so only the final mean's summation order differs)."""
import numpy as np
import pytest
from scipy import sparse

from conftest import golden


def _case():
    g = golden("evaluator.npz")
    U, I, K = (int(v) for v in g["shape"])
    mk = lambda p, i: sparse.csr_matrix((np.ones(i.shape[0]), i, p), shape=(U, I))  # noqa: E731
    return g, mk(g["train_indptr"], g["train_indices"]), mk(g["test_indptr"], g["test_indices"])


@pytest.mark.parametrize("nneg,seed,with_train", [(100, 1234, True), (100, 7, True), (20, 1234, False), (0, 3, True)])
def test_candidate_lists_bit_exact(oracle, nneg, seed, with_train):
    from cymf_b200.evaluator import Evaluator
    g, train, test = _case()
    ev = Evaluator(test, train if with_train else None, num_negatives=nneg)
    ptr, items = ev.candidates(seed)
    optr, oitems = oracle.eval_candidates(test, train if with_train else None, nneg, seed)
    assert np.array_equal(ptr, optr) and np.array_equal(items, oitems)
    users_with_items = np.diff(test.indptr) > 0
    assert np.array_equal(np.diff(ptr)[users_with_items], np.diff(test.indptr)[users_with_items] + nneg)
    assert not np.diff(ptr)[~users_with_items].any()          # users without test items get no candidates


def test_candidates_reject_all_positives(oracle):
    """Dense rows: almost every draw collides, negatives must still avoid test+train positives."""
    from cymf_b200.evaluator import Evaluator
    rng = np.random.default_rng(0)
    A = (rng.random((30, 40)) < 0.9).astype(float)
    A[:, 7] = 0                                                # at least one legal negative everywhere
    test = sparse.csr_matrix(A * (rng.random(A.shape) < 0.2))
    train = sparse.csr_matrix(A) - test
    train.eliminate_zeros()
    ev = Evaluator(test, train, num_negatives=15)
    ptr, items = ev.candidates(5)
    optr, oitems = oracle.eval_candidates(test, train, 15, 5)
    assert np.array_equal(ptr, optr) and np.array_equal(items, oitems)
    for u in range(30):
        npos = test.indptr[u + 1] - test.indptr[u]
        negs = items[ptr[u] + npos:ptr[u + 1]]
        assert not A[u, negs].any()


@pytest.mark.gpu
def test_metrics_match_reference_golden_and_oracle(oracle):
    from cymf_b200.evaluator import Evaluator, AverageOverAllEvaluator, AoaEvaluator
    assert AoaEvaluator is AverageOverAllEvaluator
    g, train, test = _case()
    for tag in "abc":
        nneg, seed, with_train = (int(v) for v in g[f"{tag}_cfg"])
        ks = [int(v) for v in g[f"{tag}_ks"]]
        ev = Evaluator(test, train if with_train else None, ["DCG", "Recall", "MAP"], ks if len(ks) > 1 else ks[0], nneg)
        res = ev.evaluate(g["W"], g["H"], seed, return_order=True)
        assert sorted(res) == [str(k) for k in g[f"{tag}_keys"]]
        got = np.array([res[str(k)] for k in g[f"{tag}_keys"]])
        assert np.abs(got - g[f"{tag}_vals"]).max() <= 1e-15
        ores, optr, oitems, oorder = oracle.evaluate(g["W"], g["H"], test, train if with_train else None, k=ks,
                                                     num_negatives=nneg, seed=seed, return_order=True)
        assert np.array_equal(ev.last_order_, oorder)              # every rank of every user, bit-exact
        for k in res:
            assert abs(res[k] - ores[k]) <= 1e-15
        res2 = ev.evaluate(g["W"], g["H"], seed)                   # cached candidate lists, same answer
        assert res2 == res


@pytest.mark.gpu
def test_ties_and_empty_users(oracle):
    """All-zero user rows (WMF zeroes users without interactions, wmf.pyx:154-156) tie every score."""
    from cymf_b200.evaluator import Evaluator
    g, train, test = _case()
    W = g["W"].copy()
    W[::3] = 0.0
    H = g["H"].copy()
    H[5] = H[9]                                                   # duplicate item vectors -> exact score ties
    ev = Evaluator(test, train, k=[1, 5])
    res = ev.evaluate(W, H, 11, return_order=True)
    ores, _, _, oorder = oracle.evaluate(W, H, test, train, k=[1, 5], seed=11, return_order=True)
    assert np.array_equal(ev.last_order_, oorder)
    for k in res:
        assert abs(res[k] - ores[k]) <= 1e-15


@pytest.mark.gpu
def test_ml100k_shape_and_k_list(oracle):
    import cymf_b200 as cymf
    train, test = cymf.synth.movielens_like("ml-100k")
    rng = np.random.default_rng(1)
    W, H = rng.normal(size=(943, 20)), rng.normal(size=(1682, 20))
    ev = cymf.evaluator.AverageOverAllEvaluator(test, train, k=5)
    res = ev.evaluate(W, H)
    ores = oracle.evaluate(W, H, test, train, k=5)
    assert set(res) == {"DCG@5", "Recall@5", "MAP@5"}
    for k in res:
        assert abs(res[k] - ores[k]) <= 1e-15
    with pytest.raises(NotImplementedError):
        cymf.evaluator.UnbiasedEvaluator(test, train).evaluate(W, H)
