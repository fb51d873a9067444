"""Probe: RelMF Hogwild stability / metrics vs. in-flight cap on the ml-100k shape (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cymf_b200 as cymf

train, test = cymf.synth.movielens_like("ml-100k")
ev = cymf.evaluator.AverageOverAllEvaluator(test, train, k=5)
for opt, lr in (("adam", 0.01), ("sgd", 0.05), ("adagrad", 0.05)):
    for inflight in (64, 256, 943, 2048, 6200, 0, None):
        m = cymf.RelMF(20, 0.1, lr, opt, 0.001, max_inflight=inflight)
        m.fit(train, 6, 8)
        fin = np.isfinite(m.W).all() and np.isfinite(m.H).all()
        r = ev.evaluate(np.nan_to_num(m.W), np.nan_to_num(m.H))
        print(opt, inflight, "finite", fin, "absmax", np.nanmax(np.abs(m.W)), {k: round(v, 4) for k, v in r.items()}, flush=True)
