"""Seeded synthetic inputs of MovieLens / co-occurrence shapes (there is no network for the real datasets).

These replace what `cymf.dataset.MovieLens(...)` / `cymf.dataset.Text8(...)` would hand to `fit()`
(cymf/dataset/movielens.py:42-86, cymf/dataset/text8.py) -- binary implicit-feedback CSR matrices with a
train / test split, and a word-word co-occurrence matrix.  The interaction matrices carry planted
cluster structure so that Recall@5 / DCG@5 / MAP@5 sit far above both the uniform-random floor (~0.045)
and the popularity-only baseline (~0.2): a BPR / WMF model has to learn the user factors to score well,
so metric regressions are visible.

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


def synth_implicit(U, I, nnz, seed, n_clusters=16, boost=40.0):
    """Binary U x I CSR with ~nnz entries (exactly nnz unless the shape is too dense to reach it):
    lognormal user activity, Zipf-like item popularity, and a per-user taste cluster that multiplies the
    odds of a 1/n_clusters slice of the catalogue by `boost`."""
    rng = np.random.default_rng(seed)
    n_clusters = int(min(n_clusters, max(2, I // 8)))
    deg = rng.lognormal(mean=0.0, sigma=1.0, size=U)
    cap = max(1, I // 4)
    deg = np.clip(np.rint(deg * (nnz / deg.sum())), 1, cap).astype(np.int64)
    for _ in range(8):                                   # redistribute what the cap removed
        short = nnz - int(deg.sum())
        room = deg < cap
        if short <= 0 or not room.any():
            break
        deg[room] = np.minimum(cap, deg[room] + np.maximum(1, short * deg[room] // max(1, int(deg[room].sum()))))
    pop = (rng.permutation(I) + 10.0) ** -0.8
    item_cluster = rng.integers(0, n_clusters, size=I)
    user_cluster = rng.integers(0, n_clusters, size=U)
    cdfs = []
    for c in range(n_clusters):
        cdf = np.cumsum(pop * np.where(item_cluster == c, boost, 1.0))
        cdfs.append(cdf / cdf[-1])
    members = [np.flatnonzero(user_cluster == c).astype(np.int64) for c in range(n_clusters)]

    def draw(counts):
        """counts[u] draws for every user from its cluster's distribution -> int64 keys u*I+i"""
        out = []
        for c in range(n_clusters):
            us = members[c]
            if us.size == 0 or counts[us].sum() == 0:
                continue
            rows = np.repeat(us, counts[us])
            cols = np.minimum(np.searchsorted(cdfs[c], rng.random(rows.shape[0])), I - 1)
            out.append(rows * I + cols)
        return np.concatenate(out) if out else np.empty(0, np.int64)

    keys = np.unique(draw(deg + deg // 3 + 2))            # oversample and dedupe ...
    for _ in range(3):                                    # ... then top up the users that fell short
        have = np.bincount(keys // I, minlength=U)
        short = np.maximum(deg - have, 0)
        if int(short.sum()) <= max(U, nnz // 200) or keys.shape[0] >= nnz:
            break
        keys = np.union1d(keys, draw(short * 3))
    if keys.shape[0] > nnz:                               # thin uniformly to ~nnz (binomial, not exact)
        keys = keys[rng.random(keys.shape[0]) < nnz / keys.shape[0]]
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


def synth_implicit_device(U, I, nnz, seed, block_users=500_000, clusters=64):
    """Device-side generator for matrices that never exist on the host (BASELINE configs[4]: 10 M x 1 M x 1 B nnz).
    Returns CSR (indptr int64, indices int32; rows sorted, deduplicated) as torch tensors on the current CUDA device:
    lognormal(sigma=1) user activity, item popularity ~ rank^-1/2, half of every user's items drawn from its taste
    cluster (u mod 64).  Identical on every rank for the same seed (torch's Philox stream)."""
    import torch

    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    mean_deg = nnz / U
    lens, idx = [], []
    for lo in range(0, U, block_users):
        n = min(block_users, U - lo)
        act = torch.exp(torch.randn(n, device=dev, generator=g))
        deg = torch.clamp(torch.round(act * (mean_deg / 1.6487) * 1.06), 1, I // 4).to(torch.int64)   # E[lognormal] = e^0.5
        rows = torch.repeat_interleave(torch.arange(lo, lo + n, device=dev), deg)
        v = torch.rand(rows.numel(), device=dev, generator=g)
        in_cluster = torch.rand(rows.numel(), device=dev, generator=g) < 0.5
        per = I // clusters
        pop_all = torch.clamp((v * v * I).to(torch.int64), max=I - 1)
        pop_clu = torch.clamp((v * v * per).to(torch.int64), max=per - 1) * clusters + rows % clusters
        items = torch.where(in_cluster, torch.clamp(pop_clu, max=I - 1), pop_all)
        keys = torch.unique(rows * I + items)                       # sorted, duplicates removed
        lens.append(torch.bincount(torch.div(keys, I, rounding_mode="floor") - lo, minlength=n))
        idx.append((keys % I).to(torch.int32))
        del keys, rows, v, items, pop_all, pop_clu, in_cluster
    lens = torch.cat(lens)
    indptr = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(lens, 0)])
    return indptr, torch.cat(idx)


