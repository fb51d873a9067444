"""Dev tool: time the WMF half sweeps under different CTA widths / staging sizes / preconditioning (one GPU).
usage: python tools/als_tune.py ml-20m 128 [width:stage_rows:solver ...]   (0 = auto; solver 0 cg, 1 pcg, 2 transformed)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import cymf_b200 as cymf
from cymf_b200.wmf import AlsSession
from cymf_b200.host import init_factors

name, K = (sys.argv[1], int(sys.argv[2])) if len(sys.argv) > 2 else ("ml-20m", 128)
train, _ = cymf.synth.movielens_like(name)
W0, H0 = init_factors(train.shape[0], train.shape[1], K)
# converge a bit first so every configuration starts from the same (epoch-3) state
base = AlsSession(train, W0.copy(), H0.copy(), 0.01, 10.0, cg_tol=1e-6, cg_max_iter=2 * K)
for _ in range(3):
    base.epoch()
Wd, Hd = base.dW.clone(), base.dH.clone()
del base
CONFIGS = [tuple(int(v) for v in c.split(":")) for c in sys.argv[3:]] or [(0, 0, 2), (0, 0, 1), (4, 0, 2), (8, 0, 2), (16, 0, 2)]
for width, stage, pre in CONFIGS:
    s = AlsSession(train, W0.copy(), H0.copy(), 0.01, 10.0, cg_tol=1e-6, cg_max_iter=2 * K, force_width=width,
                   stage_rows=stage, solver=('cg', 'pcg', 'transformed')[pre])
    res = []
    for rep in range(2):
        s.dW.copy_(Wd); s.dH.copy_(Hd)
        i0 = s.stats()[0]
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(); s.user_half(); e1.record(); s.item_half(); e2.record(); torch.cuda.synchronize()
        res = (e0.elapsed_time(e1), e1.elapsed_time(e2), s.stats()[0] - i0)
    rows = train.shape[0] + train.shape[1]
    print(f"{name} K={K} width={width or 'auto':>4} stage_rows={stage or 'auto':>4} solver={('cg', 'pcg', 'transformed')[pre]}: user {res[0]:7.2f} ms  "
          f"item {res[1]:7.2f} ms  classes u{s.classes_u} i{s.classes_i}  it/row {res[2]/rows:5.1f}", flush=True)
    del s
