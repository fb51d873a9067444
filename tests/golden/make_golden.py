#!/usr/bin/env python
"""Generate tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref, see oracle/build_ref.py).

Run in the build container only (needs /root/reference to have been compiled):
    python oracle/build_ref.py && python tests/golden/make_golden.py
The vectors pin oracle/cymf_oracle.c (tests/test_oracle_golden.py); the reference's own test-suite
holds no numeric vectors for this path (tests/test_dataset.py is shape-only and needs the network).

Every case drives the reference through the typed boundary methods named in SURVEY.md section 8(b)
(`BPR._fit_bpr`, `WMF._als` via `WMF.fit`, `GloVe._fit_glove`, `Evaluator.evaluate`, `cymf.metrics.*`)
with inputs recorded in the fixture, so the test does not depend on sklearn's / NumPy's shuffles.
The RNG vector comes from libstdc++ itself (std::mt19937 + uniform_int_distribution<long>), the
third-party dependency behind cymf/math.pyx:12-18.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np
from scipy import sparse

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
sys.path.insert(1, ROOT)

import cymf  # noqa: E402  (the compiled reference)
from cymf import metrics as ref_metrics  # noqa: E402
from cymf_b200.synth import synth_implicit, split_train_test, synth_cooc  # noqa: E402


def rng_vectors():
    src = r"""
#include <random>
#include <cstdio>
int main(){ std::mt19937 g(1234); for(int t=0;t<16;++t) printf("%u ", (unsigned)g()); printf("\n");
 long ns[4]={1682,26744,7,1000003};
 for(int q=0;q<4;++q){ std::mt19937 r(1234); std::uniform_int_distribution<long> d(0,ns[q]-1);
  for(int t=0;t<64;++t) printf("%ld ", d(r)); printf("\n"); }
 std::mt19937 r(99); std::uniform_int_distribution<long> d(0,2); for(int t=0;t<200;++t) printf("%ld ", d(r)); printf("\n");
 return 0; }
"""
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "k.cpp")
        open(p, "w").write(src)
        subprocess.check_call(["/usr/bin/g++", "-O1", p, "-o", os.path.join(d, "k")])
        lines = subprocess.check_output([os.path.join(d, "k")]).decode().strip().split("\n")
    out = {"u32_seed1234": np.array(lines[0].split(), dtype=np.uint32)}
    for q, n in enumerate((1682, 26744, 7, 1000003)):
        out[f"below_{n}_seed1234"] = np.array(lines[1 + q].split(), dtype=np.int32)
    out["below_3_seed99"] = np.array(lines[5].split(), dtype=np.int32)
    return out


def rng64_vectors():
    """uniform_int_distribution<long> over ranges at and beyond the 32-bit engine's (RelMF draws from [0, U*I))."""
    ns = (3703857792, 4294967296, 4294967297, 5000000000, 10000000000000)
    src = r"""
#include <random>
#include <cstdio>
int main(){ long ns[5]={3703857792L,4294967296L,4294967297L,5000000000L,10000000000000L};
 for(int q=0;q<5;++q){ std::mt19937 r(1234); std::uniform_int_distribution<long> d(0,ns[q]-1);
  for(int t=0;t<96;++t) printf("%ld ", d(r)); printf("\n"); }
 return 0; }
"""
    with tempfile.TemporaryDirectory() as d:
        p = os.path.join(d, "k.cpp")
        open(p, "w").write(src)
        subprocess.check_call(["/usr/bin/g++", "-O1", p, "-o", os.path.join(d, "k")])
        lines = subprocess.check_output([os.path.join(d, "k")]).decode().strip().split("\n")
    return {f"below_{n}_seed1234": np.array(lines[q].split(), dtype=np.int64) for q, n in enumerate(ns)}


def relmf_case(U, I, nnz, K, epochs, lr, wd, clip, opt, seed, ratings=False):
    """`RelMF.fit` of the compiled reference on a dense matrix (relmf.pyx:68-105 densifies sparse input anyway)."""
    X = synth_implicit(U, I, nnz, seed).astype(np.float64)
    if ratings:                                             # X[u,i] is used as a value, not only as a flag
        X.data[:] = np.random.default_rng(seed).integers(1, 6, X.nnz) / 5.0
    m = cymf.RelMF(K, clip, lr, opt, wd)
    m.fit(X.toarray(), epochs, 1)
    np.random.seed(4321)
    W0 = np.random.uniform(low=-0.1, high=0.1, size=(U, K)) / K
    H0 = np.random.uniform(low=-0.1, high=0.1, size=(I, K)) / K
    dense = X.toarray()
    prop = np.maximum(dense.mean(axis=0) / dense.mean(axis=0).max(), 1e-5) ** 0.5
    return dict(indptr=X.indptr.astype(np.int32), indices=X.indices.astype(np.int32), data=X.data,
                shape=np.array([U, I, K]), W0=W0, H0=H0, W=np.array(m.W), H=np.array(m.H), propensities=prop,
                epochs=epochs, lr=lr, wd=wd, clip=clip, opt=opt)


def bpr_case(U, I, nnz, K, epochs, lr, wd, opt, seed):
    from sklearn import utils
    X = synth_implicit(U, I, nnz, seed).astype(np.float64)
    np.random.seed(4321)                                   # bpr.pyx:97-101
    W0 = np.random.uniform(low=-0.1, high=0.1, size=(U, K)) / K
    H0 = np.random.uniform(low=-0.1, high=0.1, size=(I, K)) / K
    users, positives = utils.shuffle(*(X.nonzero()))       # bpr.pyx:104
    users, positives = users.astype(np.int32), positives.astype(np.int32)
    m = cymf.BPR(K, lr, opt, wd)
    m.W, m.H = W0.copy(), H0.copy()
    m.valid_evaluator, m.early_stopping = None, False    # set by fit() before _fit_bpr (bpr.pyx:89-92)
    m._fit_bpr(users, positives, X, epochs, lr, wd, 1, False)
    # cross-check that the full fit() (own prologue) lands on the same numbers
    m2 = cymf.BPR(K, lr, opt, wd)
    m2.fit(X, epochs, 1, verbose=False)
    assert np.array_equal(m2.W, m.W) and np.array_equal(m2.H, m.H)
    return dict(indptr=X.indptr.astype(np.int32), indices=X.indices.astype(np.int32), shape=np.array([U, I, K]),
                users=users, positives=positives, W0=W0, H0=H0, W=np.array(m.W), H=np.array(m.H),
                epochs=epochs, lr=lr, wd=wd, opt=opt)


def wmf_case(U, I, nnz, K, epochs, wd, weight, seed, empty_rows=True):
    X = synth_implicit(U, I, nnz, seed).tolil()
    if empty_rows:                                          # exercise wmf.pyx:154-156
        X[3, :] = 0
        X[:, 5] = 0
    X = X.tocsr().astype(np.float64)
    X.eliminate_zeros()
    m = cymf.WMF(K, wd, weight)
    snaps = []
    np.random.seed(4321)
    W0 = np.random.uniform(low=-0.1, high=0.1, size=(U, K)) / K
    H0 = np.random.uniform(low=-0.1, high=0.1, size=(I, K)) / K
    m.W, m.H = W0.copy(), H0.copy()
    for _ in range(epochs):
        m.fit(X, 1, 1, verbose=False)                       # warm start: one epoch per call
        snaps.append((np.array(m.W).copy(), np.array(m.H).copy()))
    m2 = cymf.WMF(K, wd, weight)
    m2.fit(X, epochs, 2, verbose=False)
    assert np.array_equal(m2.W, m.W) and np.array_equal(m2.H, m.H)
    return dict(indptr=X.indptr.astype(np.int32), indices=X.indices.astype(np.int32), shape=np.array([U, I, K]),
                W0=W0, H0=H0, W_e1=snaps[0][0], H_e1=snaps[0][1], W=snaps[-1][0], H=snaps[-1][1],
                epochs=epochs, wd=wd, weight=weight)


def glove_case(V, nnz, K, epochs, lr, x_max, alpha, seed):
    X = synth_cooc(V, nnz, seed)
    rng = np.random.default_rng(seed)
    central, context = X.nonzero()
    counts = X.data.copy()
    perm = rng.permutation(central.shape[0])
    central, context, counts = central[perm].astype(np.int32), context[perm].astype(np.int32), counts[perm]
    W0 = rng.uniform(-0.5, 0.5, size=(V, K)) / K
    H0 = rng.uniform(-0.5, 0.5, size=(V, K)) / K
    bw0 = rng.uniform(-0.5, 0.5, size=V) / K
    bh0 = rng.uniform(-0.5, 0.5, size=V) / K
    W, H, bw, bh = W0.copy(), H0.copy(), bw0.copy(), bh0.copy()
    g = cymf.GloVe(K, lr, alpha, x_max)
    g._fit_glove(central, context, counts, W, bw, H, bh, epochs, lr, x_max, alpha, 1, False)
    return dict(central=central, context=context, counts=counts, W0=W0, H0=H0, bw0=bw0, bh0=bh0,
                W=W, H=H, bw=bw, bh=bh, epochs=epochs, lr=lr, x_max=x_max, alpha=alpha)


def eval_case(U, I, nnz, K, seed):
    X = synth_implicit(U, I, nnz, seed)
    train, test = split_train_test(X, seed)
    rng = np.random.default_rng(seed)
    W = rng.normal(size=(U, K))
    H = rng.normal(size=(I, K))
    out = dict(train_indptr=train.indptr.astype(np.int32), train_indices=train.indices.astype(np.int32),
               test_indptr=test.indptr.astype(np.int32), test_indices=test.indices.astype(np.int32),
               shape=np.array([U, I, K]), W=W, H=H)
    for tag, ks, nneg, sd, with_train in (("a", 5, 100, 1234, True), ("b", [1, 5, 10], 100, 7, True),
                                          ("c", 5, 20, 1234, False)):
        ev = cymf.Evaluator(test, train if with_train else None, ["DCG", "Recall", "MAP"], ks, nneg)
        res = ev.evaluate(W, H, sd)
        keys = sorted(res)
        out[f"{tag}_keys"] = np.array(keys)
        out[f"{tag}_vals"] = np.array([res[k] for k in keys])
        out[f"{tag}_cfg"] = np.array([nneg, sd, int(with_train)])
        out[f"{tag}_ks"] = np.array([ks] if isinstance(ks, int) else ks)
    return out


def write_corpus(path, n_words, vocab, seed, lines):
    """Zipf-distributed pseudo text; with lines > 1, line boundaries are placed so that the words next to a newline
    also occur inside lines (the reference counts words on the text with newlines glued as "<eos>", glove.pyx:199-200,
    and would raise KeyError otherwise)."""
    rng = np.random.default_rng(seed)
    p = 1.0 / (np.arange(vocab) + 1.0)
    ids = rng.choice(vocab, size=n_words, p=p / p.sum())
    words = [f"w{i}" for i in ids]
    per = n_words // lines
    chunks = [words[q * per:(q + 1) * per] for q in range(lines)]
    open(path, "w").write("\n".join(" ".join(c) for c in chunks))


def cooc_case(fname, min_count, window):
    """`cymf.glove.read_text` of the compiled reference on the committed corpus file."""
    from cymf.glove import read_text
    X, i2w = read_text(os.path.join(HERE, fname), min_count, window)
    X = X.tocsr()
    X.sum_duplicates()
    X.sort_indices()
    return dict(indptr=X.indptr.astype(np.int64), indices=X.indices.astype(np.int32), data=X.data,
                shape=np.array(X.shape), words=np.array([i2w[i] for i in range(len(i2w))]),
                min_count=min_count, window=window)


def metric_vectors():
    rng = np.random.default_rng(5)
    ys, vals = [], []
    for n in (1, 2, 5, 7, 105, 130):
        for dens in (0.0, 0.1, 0.5, 1.0):
            y = (rng.random(n) < dens).astype(np.int32)
            for k in (1, 3, 5, 10):
                ys.append(np.concatenate([[n, k], y]))
                vals.append([ref_metrics.dcg_at_k(y, k), ref_metrics.recall_at_k(y, k),
                             ref_metrics.average_precision_at_k(y, k)])
    width = max(len(r) for r in ys)
    table = np.full((len(ys), width), -1, np.int32)
    for r, row in enumerate(ys):
        table[r, :len(row)] = row
    return dict(cases=table, values=np.array(vals))


def main():
    only = set(sys.argv[1:])                                # e.g. `make_golden.py relmf rng64` regenerates a subset

    def save(name, make):
        if not only or any(name.startswith(o) for o in only):
            np.savez_compressed(os.path.join(HERE, name), **make())

    if not only or "cooc" in only:
        write_corpus(os.path.join(HERE, "corpus_one_line.txt"), 6000, 300, seed=61, lines=1)      # text8-like
        write_corpus(os.path.join(HERE, "corpus_lines.txt"), 4000, 60, seed=62, lines=5)
    save("cooc_one_line.npz", lambda: cooc_case("corpus_one_line.txt", 5, 10))
    save("cooc_lines.npz", lambda: cooc_case("corpus_lines.txt", 3, 4))
    save("rng64.npz", rng64_vectors)
    for opt in ("sgd", "adagrad", "adam"):
        save(f"relmf_{opt}.npz", lambda: relmf_case(30, 40, 300, 12, 2, 0.05, 0.01, 0.1, opt, seed=51))
    save("relmf_ratings.npz", lambda: relmf_case(25, 35, 260, 20, 2, 0.001, 0.01, 0.3, "adam", seed=52, ratings=True))
    save("rng.npz", rng_vectors)
    for opt in ("sgd", "adagrad", "adam"):
        save(f"bpr_{opt}.npz", lambda: bpr_case(60, 90, 700, 20, 3, 0.01, 0.01, opt, seed=11))
    save("bpr_sgd_mid.npz", lambda: bpr_case(300, 500, 12000, 16, 4, 0.05, 0.002, "sgd", seed=12))
    save("wmf_small.npz", lambda: wmf_case(60, 90, 700, 8, 3, 0.01, 10.0, seed=21))
    save("wmf_k64.npz", lambda: wmf_case(150, 120, 3000, 64, 2, 0.01, 10.0, seed=22))
    save("glove.npz", lambda: glove_case(50, 400, 16, 3, 0.05, 10.0, 0.75, seed=31))
    save("evaluator.npz", lambda: eval_case(120, 200, 4000, 12, seed=41))
    save("metrics.npz", metric_vectors)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
