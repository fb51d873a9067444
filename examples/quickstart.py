"""End-to-end tour of cymf_b200 on synthetic data (needs a B200): implicit-feedback models with per-epoch validation
and early stopping, then the GloVe pipeline from a text file.

    python examples/quickstart.py                  # BPR, WMF and RelMF on an ml-100k-shaped matrix
    python examples/quickstart.py --models wmf --shape ml-1m --components 64
"""
import argparse
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import cymf_b200 as cymf  # noqa: E402


def build(name, K, lr, wd):
    if name == "bpr":
        return cymf.BPR(K, lr, "adam", wd)
    if name == "wmf":
        return cymf.WMF(K, wd, 10.0)
    return cymf.RelMF(K, 0.1, lr, "adam", wd)


def implicit_models(opts):
    data = cymf.dataset.SyntheticMovieLens(opts.shape)
    valid = cymf.evaluator.AverageOverAllEvaluator(data.valid, data.train, metrics=["DCG"], k=5)
    test = cymf.evaluator.AverageOverAllEvaluator(data.test, data.train, k=5)
    for name in opts.models:
        model = build(name, opts.components, opts.lr, opts.wd)
        started = time.perf_counter()
        model.fit(data.train, opts.epochs, 1, valid_evaluator=valid, early_stopping=True, verbose=False)
        took = time.perf_counter() - started
        scores = {k: round(float(v), 4) for k, v in test.evaluate(model.W, model.H).items()}
        print(f"{name:6s} fit {took:6.2f} s  best valid DCG@5 {model.valid_dcg:.4f}  test {scores}")


def word_vectors(opts):
    rng = np.random.default_rng(0)
    p = 1.0 / (np.arange(3000) + 1.0)
    ids = rng.choice(3000, size=300_000, p=p / p.sum())
    with tempfile.TemporaryDirectory() as tmp:
        corpus = os.path.join(tmp, "corpus.txt")
        with open(corpus, "w") as f:
            f.write(" ".join(f"tok{i}" for i in ids))
        X, i2w = cymf.glove.read_text(corpus, min_count=5, window_size=10)
        model = cymf.GloVe(50, 0.05)
        model.fit(X, 10, 1)
        out = os.path.join(tmp, "vectors.txt")
        model.save_word2vec_format(out, i2w)
        print(f"glove  {X.shape[0]} words, {X.nnz} co-occurrence cells, vectors written: {os.path.getsize(out)} bytes")


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--models", nargs="+", default=["bpr", "wmf", "relmf"], choices=["bpr", "wmf", "relmf"])
    ap.add_argument("--shape", default="ml-100k", choices=sorted(cymf.synth.CONFIGS))
    ap.add_argument("--components", type=int, default=20)
    ap.add_argument("--epochs", type=int, default=30)
    ap.add_argument("--lr", type=float, default=0.01)
    ap.add_argument("--wd", type=float, default=0.01)
    ap.add_argument("--skip-glove", action="store_true")
    options = ap.parse_args()
    implicit_models(options)
    if not options.skip_glove:
        word_vectors(options)
