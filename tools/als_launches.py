"""Dev tool: one eagerly launched WMF ALS epoch between cudaProfilerStart/Stop, for the ncu launch list
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv \
        python tools/als_launches.py ml-20m 128
(per-launch times under ncu are cold-cache and serialised: compare shares, not absolutes)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import cymf_b200 as cymf
from cymf_b200.wmf import AlsSession
from cymf_b200.host import init_factors

name, K = (sys.argv[1], int(sys.argv[2])) if len(sys.argv) > 2 else ("ml-20m", 128)
train, _ = cymf.synth.movielens_like(name)
W, H = init_factors(train.shape[0], train.shape[1], K)
s = AlsSession(train, W, H, 0.01, 10.0, cg_tol=1e-6, cg_max_iter=2 * K)
s.use_graph = False
for _ in range(3):
    s.epoch()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
e0.record()
s.epoch()
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"{name} K={K}: eager epoch {e0.elapsed_time(e1):.3f} ms, row solver {s.row_solver}, short_max {s.short_max}")
