"""Dev tool: WMF ALS epoch time on one GPU (not the bench)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cymf_b200 as cymf
from cymf_b200.wmf import AlsSession
from cymf_b200.host import init_factors

for name, K, dtype in [(a, int(b), c) for a, b, c in (x.split(":") for x in (sys.argv[1:] or ["ml-1m:64:float32", "ml-20m:128:float32"]))]:
    t = time.time(); train, _ = cymf.synth.movielens_like(name); print(name, "gen", round(time.time() - t, 1), train.nnz, flush=True)
    W, H = init_factors(train.shape[0], train.shape[1], K)
    t = time.time()
    s = AlsSession(train, W, H, 0.01, 10.0, dtype=dtype, cg_tol=1e-6 if dtype == "float32" else 1e-10, cg_max_iter=2 * K)
    torch.cuda.synchronize(); print("session setup", round(time.time() - t, 2), "s", flush=True)
    rows = train.shape[0] + train.shape[1]
    for e in range(6):
        i0 = s.stats()[0]
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(); s.user_half(); e1.record(); s.item_half(); e2.record(); torch.cuda.synchronize()
        it = s.stats()[0] - i0
        ms = e0.elapsed_time(e2)
        print(f"{name} K={K} {dtype} epoch {e}: {ms:8.2f} ms (user {e0.elapsed_time(e1):7.2f} item {e1.elapsed_time(e2):7.2f})  "
              f"CG it/row {it/rows:5.1f}  unconverged {s.stats()[1]}  algorithmic {s.bytes_per_epoch/ms/1e6:7.1f} GB/s", flush=True)
    s.epochs_done = 6
    del s
