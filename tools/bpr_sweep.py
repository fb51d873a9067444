"""Dev tool: BPR Hogwild epoch throughput over kernel variants on one GPU (not the bench)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from sklearn import utils
import cymf_b200 as cymf
from cymf_b200.bpr import BprSession

name = sys.argv[1] if len(sys.argv) > 1 else "ml-20m"
t = time.time(); train, test = cymf.synth.movielens_like(name); print("gen", round(time.time() - t, 1), train.nnz, flush=True)
U, I = train.shape
np.random.seed(0)
users, positives = utils.shuffle(*train.nonzero())
users, positives = users.astype(np.int32), positives.astype(np.int32)
for K in (64, 128):
    for opt, dtype, scatter in (("sgd", "float32", "store"), ("sgd", "float32", "red"), ("adam", "float32", "store"),
                                ("adagrad", "float32", "store"), ("sgd", "float64", "store"), ("sgd", "float64", "red")):
        W = np.random.uniform(-0.1, 0.1, (U, K)) / K; H = np.random.uniform(-0.1, 0.1, (I, K)) / K
        s = BprSession(W, H, users, positives, train, opt, dtype=dtype, scatter=scatter)
        for _ in range(3): s.epoch(0.01, 0.01)
        torch.cuda.synchronize()
        a0 = s.applied()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 10
        e0.record()
        for _ in range(n): s.epoch(0.01, 0.01)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        app = (s.applied() - a0) / n
        gbs = app * s.bytes_per_update / (ms * 1e-3) / 1e9
        print(f"K={K:4d} {opt:8s} {dtype:8s} {scatter:6s} {ms:8.3f} ms/epoch  {app/(ms*1e-3)/1e9:7.3f} G upd/s  accept {app/s.N:.4f}  {gbs:8.1f} GB/s algorithmic", flush=True)
        del s
