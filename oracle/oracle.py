"""ctypes front-end of oracle/liboracle.so -- TEST INFRASTRUCTURE ONLY.

Each function mirrors the reference call it restates (file:line under /root/reference):
  bpr_fit      <- cymf/bpr.pyx:117-171  (`BPR._fit_bpr`, num_threads=1)
  als_half     <- cymf/wmf.pyx:136-174  (`WMF._als`)
  glove_fit    <- cymf/glove.pyx:117-156 (`GloVe._fit_glove`)
  read_text    <- cymf/glove.pyx:183-241 (co-occurrence builder)
  relmf_fit    <- cymf/relmf.pyx:107-148 (`RelMF._fit_relmf`, num_threads=1)
  evaluate     <- cymf/evaluator.pyx:57-139 (`Evaluator.evaluate`, unbiased=False)
  rng_*        <- cymf/math.pyx:12-18 (`UniformGenerator`)
plus the Python prologues of `fit()` (seeded init + one-time shuffle), which are host logic shared
verbatim with the product (they are NumPy/sklearn calls in the reference too).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "cymf_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.check_call([cc, "-O2", "-std=c99", "-fPIC", "-shared", "-ffp-contract=off",
                               src, "-o", _SO, "-lm"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_rng_new.restype = C.c_void_p
        _lib.oracle_rng_new.argtypes = [C.c_uint32]
        _lib.oracle_rng_free.argtypes = [C.c_void_p]
        _lib.oracle_rng_fill_u32.argtypes = [C.c_void_p, C.c_void_p, C.c_int64]
        _lib.oracle_rng_fill_below.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_int64]
        _lib.oracle_bpr_fit.restype = C.c_int
        _lib.oracle_bpr_fit.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                        C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                        C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_uint32,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.oracle_rng_fill_below64.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_int64]
        _lib.oracle_relmf_fit.restype = C.c_int
        _lib.oracle_relmf_fit.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                          C.c_int32, C.c_double, C.c_double, C.c_double, C.c_int32, C.c_uint32,
                                          C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.oracle_als_half.restype = C.c_int
        _lib.oracle_als_half.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_int64, C.c_int64, C.c_int32, C.c_double, C.c_double]
        _lib.oracle_glove_fit.restype = C.c_int
        _lib.oracle_glove_fit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_int64, C.c_int64, C.c_int32, C.c_int32,
                                          C.c_double, C.c_double, C.c_double, C.c_void_p]
        _lib.oracle_cooc_count.restype = C.c_int64
        _lib.oracle_cooc_count.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                           C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
        _lib.oracle_eval_candidates.restype = C.c_int64
        _lib.oracle_eval_candidates.argtypes = [C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                                C.c_void_p, C.c_int32, C.c_uint32, C.c_void_p, C.c_void_p]
        _lib.oracle_eval_rank.restype = C.c_int
        _lib.oracle_eval_rank.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                          C.c_void_p]
        for name in ("oracle_dcg_at_k", "oracle_recall_at_k", "oracle_ap_at_k"):
            f = getattr(_lib, name)
            f.restype = C.c_double
            f.argtypes = [C.c_void_p, C.c_int64, C.c_int32]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


OPTIMIZERS = {"sgd": 0, "adagrad": 1, "adam": 2}


class Rng:
    """std::mt19937(seed) + uniform_int_distribution<long>(0, n-1) of libstdc++ >= 11."""

    def __init__(self, seed=1234):
        self._h = lib().oracle_rng_new(seed)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_rng_free(self._h)
            self._h = None

    def u32(self, count):
        out = np.empty(count, np.uint32)
        lib().oracle_rng_fill_u32(self._h, _p(out), count)
        return out

    def below(self, n, count):
        out = np.empty(count, np.int32)
        lib().oracle_rng_fill_below(self._h, n, _p(out), count)
        return out

    def below64(self, n, count):
        """any n >= 1, including ranges wider than the 32-bit engine (RelMF cells, cymf/relmf.pyx:127)"""
        out = np.empty(count, np.int64)
        lib().oracle_rng_fill_below64(self._h, n, _p(out), count)
        return out


def init_factors(U, I, K):
    """`fit()` prologue, cymf/bpr.pyx:97-101 == cymf/wmf.pyx:88-92 (both factors absent)."""
    np.random.seed(4321)
    W = np.random.uniform(low=-0.1, high=0.1, size=(U, K)) / K
    H = np.random.uniform(low=-0.1, high=0.1, size=(I, K)) / K
    return W, H


def bpr_prologue(X, K):
    """cymf/bpr.pyx:81-104: CSR coercion, seeded init, one-time shuffle of the (user, positive) pairs."""
    from sklearn import utils
    X = X.tocsr().astype(np.float64)
    W, H = init_factors(X.shape[0], X.shape[1], K)
    users, positives = utils.shuffle(*(X.nonzero()))
    return X, W, H, users.astype(np.int32), positives.astype(np.int32)


def bpr_fit(W, H, users, positives, X, num_epochs, lr, wd, optimizer="sgd", seed=1234,
            record=False, loss=False, negatives=None):
    """In-place on W, H (float64 C-contiguous). Returns dict with optional negatives/applied/loss."""
    assert W.dtype == np.float64 and H.dtype == np.float64 and W.flags.c_contiguous and H.flags.c_contiguous
    X = X.tocsr()
    X.sort_indices()
    indptr = np.ascontiguousarray(X.indptr, np.int32)
    indices = np.ascontiguousarray(X.indices, np.int32)
    users = np.ascontiguousarray(users, np.int32)
    positives = np.ascontiguousarray(positives, np.int32)
    N = users.shape[0]
    neg = np.empty(num_epochs * N, np.int32) if record else None
    app = np.empty(num_epochs * N, np.uint8) if record else None
    ls = np.empty(num_epochs, np.float64) if loss else None
    if negatives is not None:
        negatives = np.ascontiguousarray(negatives, np.int32)
        assert negatives.shape[0] == num_epochs * N
    rc = lib().oracle_bpr_fit(_p(W), _p(H), X.shape[0], X.shape[1], W.shape[1], _p(users), _p(positives), N,
                              _p(indptr), _p(indices), num_epochs, lr, wd, OPTIMIZERS[optimizer], seed,
                              _p(negatives), _p(neg), _p(app), _p(ls))
    if rc:
        raise MemoryError("oracle_bpr_fit")
    return {"negatives": neg, "applied": app, "loss": ls}


def relmf_propensities(X):
    """cymf/relmf.pyx:90: sqrt of the relative item popularity, floored at 1e-5 (dense NumPy in the reference)."""
    X = np.asarray(X.todense() if hasattr(X, "todense") else X, dtype=np.float64)
    return np.maximum(X.mean(axis=0) / X.mean(axis=0).max(), 1e-5) ** 0.5


def relmf_fit(W, H, X, num_epochs, lr, wd, clip, optimizer="adam", seed=1234, propensities=None,
              cells=None, record=False, loss=False, n_samples=None):
    """`RelMF._fit_relmf` (cymf/relmf.pyx:107-148), num_threads=1; in place on W, H."""
    from scipy import sparse
    assert W.dtype == np.float64 and H.dtype == np.float64 and W.flags.c_contiguous and H.flags.c_contiguous
    Xs = sparse.csr_matrix(X).astype(np.float64)
    Xs.sum_duplicates()
    Xs.sort_indices()
    U, I = Xs.shape
    if propensities is None:
        propensities = relmf_propensities(Xs)
    propensities = np.ascontiguousarray(np.asarray(propensities).ravel(), np.float64)
    n = int(U) * int(I) if n_samples is None else int(n_samples)
    indptr = np.ascontiguousarray(Xs.indptr, np.int32)
    indices = np.ascontiguousarray(Xs.indices, np.int32)
    data = np.ascontiguousarray(Xs.data, np.float64)
    rec = np.empty(num_epochs * n, np.int64) if record else None
    ls = np.empty(num_epochs, np.float64) if loss else None
    if cells is not None:
        cells = np.ascontiguousarray(cells, np.int64)
        assert cells.shape[0] == num_epochs * n
    rc = lib().oracle_relmf_fit(_p(W), _p(H), U, I, W.shape[1], _p(indptr), _p(indices), _p(data), _p(propensities), n,
                                num_epochs, lr, wd, clip, OPTIMIZERS[optimizer], seed, _p(cells), _p(rec), _p(ls))
    if rc:
        raise MemoryError("oracle_relmf_fit")
    return {"cells": rec, "loss": ls}


def als_half(indptr, indices, X, Y, wd, weight):
    """`WMF._als(indptr, indices, X, Y, num_threads)`: solves every row of X in place."""
    assert X.dtype == np.float64 and Y.dtype == np.float64 and X.flags.c_contiguous and Y.flags.c_contiguous
    indptr = np.ascontiguousarray(indptr, np.int64)
    indices = np.ascontiguousarray(indices, np.int32)
    rc = lib().oracle_als_half(_p(indptr), _p(indices), _p(X), _p(Y), X.shape[0], Y.shape[0], X.shape[1],
                               wd, weight)
    if rc:
        raise MemoryError("oracle_als_half")


def wmf_fit(X, K, wd, weight, num_epochs, W=None, H=None):
    """`WMF.fit` (cymf/wmf.pyx:59-132) without evaluator."""
    X = X.tocsr().astype(np.float64)
    if W is None or H is None:
        W, H = init_factors(X.shape[0], X.shape[1], K)
    XT = X.T.tocsr()
    for _ in range(num_epochs):
        als_half(X.indptr, X.indices, W, H, wd, weight)
        als_half(XT.indptr, XT.indices, H, W, wd, weight)
    return W, H


def glove_fit(central, context, counts, W, bw, H, bh, num_epochs, lr, x_max, alpha, loss=False):
    """`GloVe._fit_glove` with caller-supplied arrays; in place on W, bw, H, bh."""
    for a in (W, bw, H, bh):
        assert a.dtype == np.float64 and a.flags.c_contiguous
    central = np.ascontiguousarray(central, np.int32)
    context = np.ascontiguousarray(context, np.int32)
    counts = np.ascontiguousarray(counts, np.float64)
    assert bw.shape[0] >= W.shape[0] and bh.shape[0] >= H.shape[0] or True
    ls = np.empty(num_epochs, np.float64) if loss else None
    rc = lib().oracle_glove_fit(_p(central), _p(context), _p(counts), central.shape[0], _p(W), _p(bw), _p(H), _p(bh),
                                bw.shape[0], bh.shape[0], W.shape[1], num_epochs, lr, x_max, alpha, _p(ls))
    if rc:
        raise MemoryError("oracle_glove_fit")
    return ls


def read_text_vocabulary(fname, min_count=5):
    """The Python part of `read_text` (cymf/glove.pyx:198-214), statement for statement: word counts over the text with
    newlines glued as "<eos>", ids in first-appearance order of the words kept.  Returns (list of id lists, i2w)."""
    from collections import Counter
    with open(fname) as f:
        raw = f.read()
        words = raw.replace("\n", "<eos>").split(" ")
    count = dict(Counter(words))
    lines = raw.split("\n")
    w2i, i2w, x = {}, {}, []
    for i in range(len(lines)):
        words = lines[i].split(" ")
        tmp = []
        for j in range(len(words)):
            if words[j] not in w2i and count[words[j]] >= min_count:
                index = len(w2i)
                w2i[words[j]] = index
                i2w[index] = words[j]
                tmp.append(index)
            elif count[words[j]] >= min_count:
                tmp.append(w2i[words[j]])
        x.append(tmp)
    return x, i2w


def cooc_count(x, vocab_size, window_size):
    """The counting loop of `read_text` (cymf/glove.pyx:218-221) -> (rows, cols, vals) sorted by (row, col)."""
    tokens = np.ascontiguousarray(np.concatenate([np.asarray(t, np.int32) for t in x] + [np.empty(0, np.int32)]), np.int32)
    line_ptr = np.concatenate([[0], np.cumsum([len(t) for t in x])]).astype(np.int64)
    cap = int(tokens.shape[0]) * int(window_size) + 1
    rows, cols, vals = np.empty(cap, np.int32), np.empty(cap, np.int32), np.empty(cap, np.float64)
    m = lib().oracle_cooc_count(_p(tokens), _p(line_ptr), len(x), vocab_size, window_size, _p(rows), _p(cols), _p(vals), cap)
    if m < 0:
        raise MemoryError("oracle_cooc_count")
    return rows[:m].copy(), cols[:m].copy(), vals[:m].copy()


def read_text(fname, min_count=5, window_size=10):
    """`cymf.glove.read_text(fname, min_count, window_size)` (cymf/glove.pyx:183-241) -> (csr_matrix, i2w)."""
    from scipy import sparse
    x, i2w = read_text_vocabulary(fname, min_count)
    V = len(i2w)
    rows, cols, vals = cooc_count(x, max(V, 1), window_size)
    return sparse.csr_matrix((vals, (rows, cols)), shape=(V, V)), i2w


def eval_candidates(test, train, num_negatives=100, seed=1234):
    """Candidate lists of `Evaluator.evaluate` (cymf/evaluator.pyx:95-111). Returns (cand_ptr, cand_items)."""
    from scipy import sparse
    test = sparse.csr_matrix(test)
    allp = test.copy()
    if train is not None:
        allp = allp + sparse.csr_matrix(train)
    allp = allp.tocsr()
    allp.sort_indices()
    U, I = test.shape
    tip = np.ascontiguousarray(test.indptr, np.int32)
    tix = np.ascontiguousarray(test.indices, np.int32)
    aip = np.ascontiguousarray(allp.indptr, np.int32)
    aix = np.ascontiguousarray(allp.indices, np.int32)
    n_eval = int((np.diff(tip) > 0).sum())
    cand_ptr = np.empty(U + 1, np.int64)
    cand_items = np.empty(test.nnz + n_eval * num_negatives, np.int32)
    n = lib().oracle_eval_candidates(U, I, _p(tip), _p(tix), _p(aip), _p(aix), num_negatives, seed,
                                     _p(cand_ptr), _p(cand_items))
    assert n == cand_items.shape[0]
    return cand_ptr, cand_items


def evaluate(W, H, test, train=None, k=5, num_negatives=100, seed=1234, metrics=("DCG", "Recall", "MAP"),
             return_order=False):
    """`Evaluator(test, train, metrics, k, num_negatives).evaluate(W, H, seed)` -> {"DCG@5": ...}."""
    from scipy import sparse
    test = sparse.csr_matrix(test)
    W = np.ascontiguousarray(W, np.float64)
    H = np.ascontiguousarray(H, np.float64)
    cand_ptr, cand_items = eval_candidates(test, train, num_negatives, seed)
    ks = np.ascontiguousarray([k] if isinstance(k, int) else list(k), np.int32)
    out = np.zeros((ks.shape[0], 3), np.float64)
    order = np.empty(cand_items.shape[0], np.int32) if return_order else None
    tip = np.ascontiguousarray(test.indptr, np.int32)
    rc = lib().oracle_eval_rank(_p(W), _p(H), test.shape[0], W.shape[1], _p(tip), _p(cand_ptr), _p(cand_items),
                                _p(ks), ks.shape[0], _p(out), _p(order))
    if rc:
        raise MemoryError("oracle_eval_rank")
    col = {"DCG": 0, "Recall": 1, "MAP": 2}
    res = {f"{m}@{int(kk)}": float(out[q, col[m]]) for q, kk in enumerate(ks) for m in metrics}
    if return_order:
        return res, cand_ptr, cand_items, order
    return res


def metric_at_k(name, y, k):
    y = np.ascontiguousarray(y, np.int32)
    f = {"DCG": lib().oracle_dcg_at_k, "Recall": lib().oracle_recall_at_k, "MAP": lib().oracle_ap_at_k}[name]
    return f(_p(y), y.shape[0], k)
