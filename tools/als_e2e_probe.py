"""Dev tool: where the time of one `WMF.fit(X, 1)` call with host buffers goes (the e2e arm of bench.py)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cymf_b200 as cymf
from cymf_b200 import wmf as wmf_mod
from cymf_b200.host import init_factors

train, _ = cymf.synth.movielens_like("ml-20m")
K = 128
W, H = init_factors(train.shape[0], train.shape[1], K)
m = cymf.WMF(K, 0.01, 10.0)
m.W, m.H = W, H
for _ in range(2):
    m.fit(train, 1, 1, verbose=False)
torch.cuda.synchronize()

def timed(label, fn):
    torch.cuda.synchronize(); t = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print(f"{label:40s} {1e3 * (time.perf_counter() - t):8.2f} ms", flush=True); return r

for rep in range(2):
    print("--- rep", rep)
    timed("fit(X, 1) total", lambda: m.fit(train, 1, 1, verbose=False))
    X = timed("X.tocsr().astype(float64, copy=False)", lambda: train.tocsr().astype(np.float64, copy=False))
    sess = timed("AlsSession(...)", lambda: wmf_mod.AlsSession(X, m.W, m.H, 0.01, 10.0, cg_tol=1e-6, cg_max_iter=2 * K))
    timed("  _prepare_on_device alone", lambda: sess._prepare_on_device(X))
    timed("  _classes x2", lambda: (sess._classes(sess.csr_u[0]), sess._classes(sess.csr_i[0])))
    timed("  _upload W,H", lambda: (sess._upload(m.W, sess.slot_u), sess._upload(m.H, sess.slot_i)))
    timed("epoch()", sess.epoch)
    timed("download", lambda: sess.download(m.W, m.H))
    timed("stats", sess.stats)
