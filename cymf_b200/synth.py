"""Seeded synthetic inputs of MovieLens / co-occurrence shapes (there is no network for the real datasets).

These replace what `cymf.dataset.MovieLens(...)` / `cymf.dataset.Text8(...)` would hand to `fit()`
(cymf/dataset/movielens.py:42-86, cymf/dataset/text8.py) -- binary implicit-feedback CSR matrices with a
train / test split, and a word-word co-occurrence matrix.  The interaction matrices carry planted
cluster structure so that Recall@5 / DCG@5 / MAP@5 are far from the uniform-random floor and metric
regressions are visible.

Shapes of BASELINE.json's configs:
    C1  943 x 1,682 x 100 k   seed 100      C2  6,040 x 3,706 x 1 M   seed 101
    C3  138,493 x 26,744 x 20 M seed 102    C4  V = 400 k, 100 M co-occurrences seed 103
"""
import numpy as np
from scipy import sparse

CONFIGS = {
    "ml-100k": dict(U=943, I=1682, nnz=100_000, seed=100),
    "ml-1m": dict(U=6040, I=3706, nnz=1_000_000, seed=101),
    "ml-20m": dict(U=138_493, I=26_744, nnz=20_000_000, seed=102),
}


def synth_implicit(U, I, nnz, seed, n_clusters=64, boost=12.0):
    """Binary U x I CSR with ~nnz entries: lognormal user activity, Zipf-like item popularity,
    and a per-user taste cluster that boosts a 1/n_clusters slice of the catalogue."""
    rng = np.random.default_rng(seed)
    n_clusters = int(min(n_clusters, max(2, I // 8)))
    deg = rng.lognormal(mean=0.0, sigma=1.0, size=U)
    deg = np.clip(np.rint(deg * (nnz * 1.08 / deg.sum())), 1, max(1, I // 4)).astype(np.int64)
    pop = (rng.permutation(I) + 10.0) ** -0.8
    item_cluster = rng.integers(0, n_clusters, size=I)
    user_cluster = rng.integers(0, n_clusters, size=U)
    rows = np.repeat(np.arange(U, dtype=np.int64), deg)
    cols = np.empty(rows.shape[0], dtype=np.int64)
    row_cluster = user_cluster[rows]
    order = np.argsort(row_cluster, kind="stable")
    bounds = np.searchsorted(row_cluster[order], np.arange(n_clusters + 1))
    for c in range(n_clusters):
        sel = order[bounds[c]:bounds[c + 1]]
        if sel.size == 0:
            continue
        p = pop * np.where(item_cluster == c, boost, 1.0)
        cdf = np.cumsum(p)
        cdf /= cdf[-1]
        cols[sel] = np.minimum(np.searchsorted(cdf, rng.random(sel.size)), I - 1)
    keys = np.unique(rows * I + cols)
    if keys.shape[0] > nnz:
        keys = np.sort(rng.choice(keys, size=nnz, replace=False))
    X = sparse.csr_matrix((np.ones(keys.shape[0]), (keys // I, keys % I)), shape=(U, I))
    X.sort_indices()
    return X


def split_train_test(X, seed, test_frac=0.1):
    """Bernoulli hold-out over the nonzeros (mirrors cymf/dataset/movielens.py:62-66)."""
    rng = np.random.default_rng(seed + 7)
    coo = X.tocoo()
    mask = rng.random(coo.nnz) < test_frac
    test = sparse.csr_matrix((coo.data[mask], (coo.row[mask], coo.col[mask])), shape=X.shape)
    train = sparse.csr_matrix((coo.data[~mask], (coo.row[~mask], coo.col[~mask])), shape=X.shape)
    train.sort_indices()
    test.sort_indices()
    return train, test


def movielens_like(name):
    """(train, test) CSR pair of one of the MovieLens shapes in CONFIGS."""
    cfg = CONFIGS[name]
    X = synth_implicit(cfg["U"], cfg["I"], cfg["nnz"], cfg["seed"])
    return split_train_test(X, cfg["seed"])


def synth_cooc(V, nnz, seed):
    """V x V co-occurrence CSR: Zipf(1.0) rows/cols, counts = max(exp(N(0, 1.5)), 0.1)."""
    rng = np.random.default_rng(seed)
    cdf = np.cumsum(1.0 / (np.arange(V) + 1.0))
    cdf /= cdf[-1]
    m = int(nnz * 1.25)
    r = np.minimum(np.searchsorted(cdf, rng.random(m)), V - 1).astype(np.int64)
    c = np.minimum(np.searchsorted(cdf, rng.random(m)), V - 1).astype(np.int64)
    keys = np.unique(r * V + c)
    if keys.shape[0] > nnz:
        keys = np.sort(rng.choice(keys, size=nnz, replace=False))
    counts = np.maximum(np.exp(rng.normal(0.0, 1.5, size=keys.shape[0])), 0.1)
    X = sparse.csr_matrix((counts, (keys // V, keys % V)), shape=(V, V))
    X.sort_indices()
    return X
