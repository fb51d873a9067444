"""Pins oracle/cymf_oracle.c against vectors produced by the compiled reference (tests/golden/make_golden.py).

BPR / RelMF / GloVe: the oracle follows the reference's operation order exactly, so equality is bitwise.
WMF: the reference's Gram is a BLAS dgemm and its solve a LAPACK dgesv (different summation order), so
the bound is 1e-10 relative -- six orders below the 1e-4 the product is held to.
Evaluator: metric means equal to 1e-15 (scores come from BLAS `np.dot` in the reference).
"""
import numpy as np
import pytest
from scipy import sparse

from conftest import golden


def test_rng_known_answers(oracle):
    g = golden("rng.npz")
    assert np.array_equal(oracle.Rng(1234).u32(16), g["u32_seed1234"])
    # SURVEY.md section 0, fact 10
    assert list(g["u32_seed1234"][:4]) == [822569775, 2137449171, 2671936806, 3512589365]
    assert list(g["below_1682_seed1234"][:10]) == [322, 837, 1046, 1375, 736, 1029, 1320, 1297, 1311, 1447]
    for n in (1682, 26744, 7, 1000003):
        assert np.array_equal(oracle.Rng(1234).below(n, 64), g[f"below_{n}_seed1234"])
    assert np.array_equal(oracle.Rng(99).below(3, 200), g["below_3_seed99"])


def test_rng_wide_ranges(oracle):
    """Ranges at / beyond the engine's 2^32 (libstdc++'s raw-word and upscaling branches): RelMF draws cells
    from [0, U*I) (cymf/relmf.pyx:127), 3.7e9 at the ml-20m shape and 1e13 at C5's."""
    g = golden("rng64.npz")
    for n in (3703857792, 4294967296, 4294967297, 5000000000, 10000000000000):
        want = g[f"below_{n}_seed1234"]
        assert np.array_equal(oracle.Rng(1234).below64(n, want.shape[0]), want), n
        assert want.max() < n and want.min() >= 0


@pytest.mark.parametrize("name", ["relmf_sgd", "relmf_adagrad", "relmf_adam", "relmf_ratings"])
def test_relmf_bitwise(oracle, name):
    g = golden(name + ".npz")
    U, I, K = g["shape"]
    X = sparse.csr_matrix((g["data"], g["indices"], g["indptr"]), shape=(U, I))
    assert np.array_equal(oracle.relmf_propensities(X), g["propensities"])
    W, H = g["W0"].copy(), g["H0"].copy()
    oracle.relmf_fit(W, H, X, int(g["epochs"]), float(g["lr"]), float(g["wd"]), float(g["clip"]), str(g["opt"]))
    assert np.array_equal(W, g["W"])
    assert np.array_equal(H, g["H"])


@pytest.mark.parametrize("name", ["bpr_sgd", "bpr_adagrad", "bpr_adam", "bpr_sgd_mid"])
def test_bpr_bitwise(oracle, name):
    g = golden(name + ".npz")
    U, I, K = g["shape"]
    X = sparse.csr_matrix((np.ones(g["indices"].shape[0]), g["indices"], g["indptr"]), shape=(U, I))
    W, H = g["W0"].copy(), g["H0"].copy()
    oracle.bpr_fit(W, H, g["users"], g["positives"], X, int(g["epochs"]), float(g["lr"]), float(g["wd"]),
                   str(g["opt"]))
    assert np.array_equal(W, g["W"])
    assert np.array_equal(H, g["H"])


def test_bpr_prologue_matches_reference(oracle):
    g = golden("bpr_sgd.npz")
    U, I, K = g["shape"]
    X = sparse.csr_matrix((np.ones(g["indices"].shape[0]), g["indices"], g["indptr"]), shape=(U, I))
    _, W0, H0, users, positives = oracle.bpr_prologue(X, int(K))
    assert np.array_equal(W0, g["W0"]) and np.array_equal(H0, g["H0"])
    assert np.array_equal(users, g["users"]) and np.array_equal(positives, g["positives"])


@pytest.mark.parametrize("name", ["wmf_small", "wmf_k64"])
def test_wmf(oracle, name):
    g = golden(name + ".npz")
    U, I, K = g["shape"]
    X = sparse.csr_matrix((np.ones(g["indices"].shape[0]), g["indices"], g["indptr"]), shape=(U, I))
    assert (np.diff(X.indptr) == 0).any() and (np.diff(X.T.tocsr().indptr) == 0).any()  # empty rows present
    W, H = oracle.wmf_fit(X, int(K), float(g["wd"]), float(g["weight"]), 1, g["W0"].copy(), g["H0"].copy())
    for got, want in ((W, g["W_e1"]), (H, g["H_e1"])):
        assert np.abs(got - want).max() <= 1e-10 * np.abs(want).max()
    W, H = oracle.wmf_fit(X, int(K), float(g["wd"]), float(g["weight"]), int(g["epochs"]) - 1, W, H)
    for got, want in ((W, g["W"]), (H, g["H"])):
        assert np.abs(got - want).max() <= 1e-10 * np.abs(want).max()
    assert not W[3].any() and not H[5].any()      # wmf.pyx:154-156: rows without interactions are zeroed


def test_glove_bitwise(oracle):
    g = golden("glove.npz")
    W, H, bw, bh = g["W0"].copy(), g["H0"].copy(), g["bw0"].copy(), g["bh0"].copy()
    oracle.glove_fit(g["central"], g["context"], g["counts"], W, bw, H, bh, int(g["epochs"]), float(g["lr"]),
                     float(g["x_max"]), float(g["alpha"]))
    for got, want in ((W, g["W"]), (H, g["H"]), (bw, g["bw"]), (bh, g["bh"])):
        assert np.array_equal(got, want)


def test_evaluator(oracle):
    g = golden("evaluator.npz")
    U, I, K = g["shape"]
    mk = lambda p, i: sparse.csr_matrix((np.ones(i.shape[0]), i, p), shape=(U, I))  # noqa: E731
    train, test = mk(g["train_indptr"], g["train_indices"]), mk(g["test_indptr"], g["test_indices"])
    for tag in "abc":
        nneg, seed, with_train = (int(v) for v in g[f"{tag}_cfg"])
        ks = [int(v) for v in g[f"{tag}_ks"]]
        res = oracle.evaluate(g["W"], g["H"], test, train if with_train else None, k=ks, num_negatives=nneg,
                              seed=seed)
        assert sorted(res) == [str(k) for k in g[f"{tag}_keys"]]
        got = np.array([res[str(k)] for k in g[f"{tag}_keys"]])
        assert np.abs(got - g[f"{tag}_vals"]).max() <= 1e-15


def test_metric_functions(oracle):
    g = golden("metrics.npz")
    for row, want in zip(g["cases"], g["values"]):
        n, k = int(row[0]), int(row[1])
        y = row[2:2 + n]
        got = [oracle.metric_at_k(m, y, k) for m in ("DCG", "Recall", "MAP")]
        assert got == list(want)


@pytest.mark.parametrize("name,txt", [("cooc_one_line", "corpus_one_line.txt"), ("cooc_lines", "corpus_lines.txt")])
def test_read_text_bitwise(oracle, name, txt):
    """`read_text` (cymf/glove.pyx:183-241): vocabulary order and every co-occurrence count, bit for bit."""
    import os
    from conftest import GOLDEN
    g = golden(name + ".npz")
    X, i2w = oracle.read_text(os.path.join(GOLDEN, txt), int(g["min_count"]), int(g["window"]))
    X.sort_indices()
    assert X.shape == tuple(g["shape"]) and [i2w[i] for i in range(len(i2w))] == list(g["words"])
    assert np.array_equal(X.indptr, g["indptr"]) and np.array_equal(X.indices, g["indices"])
    assert np.array_equal(X.data, g["data"])
