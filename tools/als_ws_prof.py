"""Dev tool: cycle accounting of the warp-specialised row solver's roles (one CTA, the middle one of the grid) from
the kernel's debug buffer.   python tools/als_ws_prof.py [ml-20m] [128]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cymf_b200 as cymf
from cymf_b200.wmf import AlsSession
from cymf_b200.host import init_factors

name, K = (sys.argv[1], int(sys.argv[2])) if len(sys.argv) > 2 else ("ml-20m", 128)
train, _ = cymf.synth.movielens_like(name)
W, H = init_factors(train.shape[0], train.shape[1], K)
s = AlsSession(train, W, H, 0.01, 10.0, cg_tol=float(os.environ.get('CYMF_PROBE_CGTOL', '1e-6')), cg_max_iter=2 * K)
s.use_graph = False
for _ in range(3):
    s.epoch()
for side in ("user", "item"):
    s.d_debug.zero_()
    (s.user_half if side == "user" else s.item_half)()
    torch.cuda.synchronize()
    d = s.d_debug.cpu().tolist()
    print(f"{side} half, middle CTA (cycles): failures {d[0]}")
    for g in (0, 1):
        print(f"  solver {g}: wait chains {d[8 + 4 * g]:>9d}  fold {d[9 + 4 * g]:>8d}  b + CG {d[10 + 4 * g]:>9d}  rows {d[11 + 4 * g]}"
              f"  -> {d[10 + 4 * g] / max(d[11 + 4 * g], 1):.0f} cycles of CG per row")
    print(f"  solver 0 CG loop: publish r {d[26]}  matvec {d[27]}  reduction {d[28]}  iterations {d[29]}"
          f"  -> {d[26] / max(d[29], 1):.0f} + {d[27] / max(d[29], 1):.0f} + {d[28] / max(d[29], 1):.0f} cycles per iteration")
    print(f"  convert 0: wait lo slot {d[16]:>9d}  wait landed {d[17]:>9d}  wait b slot {d[18]:>8d}  total {d[19]}")
    print(f"  copy 0   : wait hi slot {d[24]:>9d}  total {d[25]}")
    print(f"  mma     : wait accumulator {d[20]:>9d}  wait stage full {d[21]:>9d}  total {d[22]}  chunks {d[23]}"
          f"  -> {d[22] / max(d[23], 1):.0f} cycles per chunk (813 = tensor-bound)")
