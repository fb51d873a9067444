"""Small non-Hogwild workload for compute-sanitizer: WMF (Gram, Cholesky transforms, GEMMs, row
solvers), the evaluator and the device-side sparse preparation, each checked against the oracle / scipy.
    compute-sanitizer --tool memcheck python tools/sanitize_small.py            # tensor-core kernels included
    CYMF_NO_TCGEN05=1 compute-sanitizer --tool racecheck python tools/sanitize_small.py   # CUDA-core kernels only
(The Hogwild kernels race by design and are left out.)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cymf_b200 as cymf
from cymf_b200 import prep
from oracle import oracle

X = cymf.synth.synth_implicit(300, 220, 9000, seed=7).tolil()
X[2, :] = 1
X = X.tocsr()
# (K, dtype, environment): default = dual-form tiles for rows of <= 64 entries + warp-specialised one-pass solver (cp.async
# gather); then the same solver with the TMA tile::gather4 copy path; the CTA-per-row solver; the streaming CG kernel
for K, dtype, env in ((32, "float32", {}), (64, "float32", {"CYMF_ALS_WS_TMA": "1", "CYMF_ALS_DUAL": "0"}),
                      (64, "float32", {"CYMF_ALS_WS": "0", "CYMF_ALS_DUAL": "0", "CYMF_ALS_SHORT": "0"}), (20, "float64", {})):
    for k in ("CYMF_ALS_WS_TMA", "CYMF_ALS_DUAL", "CYMF_ALS_WS", "CYMF_ALS_SHORT"):
        os.environ.pop(k, None)
    os.environ.update(env)
    short = env
    Wo, Ho = oracle.wmf_fit(X, K, 0.01, 10.0, 2)
    m = cymf.WMF(K, 0.01, 10.0, dtype=dtype)
    m.fit(X, 2, 1, verbose=False)
    err = max(np.abs(m.W - Wo).max() / np.abs(Wo).max(), np.abs(m.H - Ho).max() / np.abs(Ho).max())
    print(f"WMF K={K} {dtype} short={short}: rel err {err:.2e}", flush=True)
    assert err < 1e-4
train, test = cymf.synth.split_train_test(X, 5)
got = cymf.evaluator.AverageOverAllEvaluator(test, train, k=5).evaluate(m.W, m.H)
want = oracle.evaluate(m.W, m.H, test, train, k=5)
assert all(abs(got[k] - want[k]) < 1e-12 for k in want)
XT = X.T.tocsr(); XT.sort_indices()
t_ip, t_ix = prep.transpose_csr(torch.from_numpy(X.indptr.astype(np.int64)).cuda(),
                                torch.from_numpy(X.indices.astype(np.int32)).cuda(), X.shape[0], X.shape[1])
assert np.array_equal(t_ip.cpu().numpy(), XT.indptr) and np.array_equal(t_ix.cpu().numpy(), XT.indices)
r = cymf.BPR(8, 0.05, "sgd", 0.01, mode="replay")
r.fit(X[:40, :60].tocsr(), num_epochs=1, verbose=False)
print("sanitize_small ok", flush=True)
