"""Hogwild parity at the benchmarked scale (BASELINE.json configs[2] shape, 138,493 x 26,744, ~18 M train pairs):
the concurrent f32 BPR kernel against the COMPILED reference (oracle/_ref, cymf/bpr.pyx:162-169 under OpenMP with
every host core), same init (seed 4321 prologue), same hyper-parameters, a few epochs each; Recall@5 / DCG@5 / MAP@5
as means over 5 evaluator seeds (optuna_example.py:63-65), scored for BOTH sides by cymf_b200's evaluator (bit-exact
against the reference's, tests/test_evaluator.py).  Bar (north star): within 1 % relative.

    python tools/hogwild_parity_c3.py [--K 64] [--epochs 3] [--opts sgd,adam]
Prints one JSON line per configuration.  tests/test_bpr_gpu.py::test_hogwild_metric_parity_c3_shape asserts on it.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

LR = {"sgd": 0.05, "adagrad": 0.05, "adam": 0.002}


def compare(K, opt, epochs, train=None, test=None, evaluator=None, lr=None, wd=0.01):
    import cymf_b200 as cymf
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    import cymf as ref                                        # the compiled reference, not cymf_b200
    if train is None:
        train, test = cymf.synth.movielens_like("ml-20m")
    ev = evaluator or cymf.evaluator.AverageOverAllEvaluator(test, train, k=5)
    lr = LR[opt] if lr is None else lr
    threads = os.cpu_count()

    def metrics(W, H):
        rs = [ev.evaluate(W, H, seed=s) for s in range(5)]
        return {k: float(np.mean([r[k] for r in rs])) for k in rs[0]}

    t0 = time.perf_counter()
    g = cymf.BPR(K, lr, opt, wd)
    g.fit(train, num_epochs=epochs, num_threads=threads, verbose=False)
    t_gpu = time.perf_counter() - t0
    got = metrics(g.W, g.H)
    t0 = time.perf_counter()
    r = ref.BPR(K, lr, opt, wd)
    r.fit(train, epochs, threads, verbose=False)
    t_ref = time.perf_counter() - t0
    want = metrics(np.asarray(r.W), np.asarray(r.H))
    rel = {k: abs(got[k] - want[k]) / want[k] for k in want}
    return {"shape": list(train.shape), "pairs": int(train.nnz), "K": K, "optimizer": opt, "lr": lr, "wd": wd,
            "epochs": epochs, "reference_threads": threads, "reference": want, "gpu": got, "rel_diff": rel,
            "max_rel_diff": max(rel.values()), "acceptance_gpu": g.n_applied_ / max(g.n_attempted_, 1),
            "fit_sec": {"gpu_incl_upload": t_gpu, "reference_incl_setup": t_ref}}


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--K", type=int, default=64)
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--opts", default="sgd,adam")
    a = ap.parse_args()
    import cymf_b200 as cymf
    train, test = cymf.synth.movielens_like("ml-20m")
    ev = cymf.evaluator.AverageOverAllEvaluator(test, train, k=5)
    for opt in a.opts.split(","):
        print(json.dumps(compare(a.K, opt, a.epochs, train, test, ev)), flush=True)
