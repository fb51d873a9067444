"""Dev tool: where a (sharded) WMF ALS epoch spends its time, phase by phase, from CUDA events on rank 0's stream
(eager launches, no CUDA graph).   [torchrun --nproc-per-node N] python tools/als_phases.py [ml-20m] [128]"""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import cymf_b200 as cymf
from cymf_b200.wmf import AlsSession
from cymf_b200.host import init_factors

name, K = (sys.argv[1], int(sys.argv[2])) if len(sys.argv) > 2 else ("ml-20m", 128)
world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
train, _ = cymf.synth.movielens_like(name)
W, H = init_factors(train.shape[0], train.shape[1], K)
s = AlsSession(train, W, H, 0.01, 10.0, cg_tol=float(os.environ.get('CYMF_PROBE_CGTOL', '1e-6')), cg_max_iter=2 * K, distributed=world > 1)
for _ in range(3):
    s.epoch()
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
agg = collections.OrderedDict()
EPOCHS = 5
tot = 0.0
for _ in range(EPOCHS):
    s.trace = []
    s.epoch()
    torch.cuda.synchronize()
    tr, s.trace = s.trace, None
    for (l0, e0), (l1, e1) in zip(tr[:-1], tr[1:]):
        agg[l1] = agg.get(l1, 0.0) + e0.elapsed_time(e1)
    tot += tr[0][1].elapsed_time(tr[-1][1])
if s.rank == 0:
    print(f"{name} K={K} world={world}: eager epoch {tot / EPOCHS:.3f} ms on rank 0")
    for k, v in agg.items():
        print(f"   {v / EPOCHS * 1e3:9.1f} us  {k}")
s.user_half()
ru = s.residual("user", 24, seed=1)
s.item_half()
ri = s.residual("item", 24, seed=2)
if s.rank == 0:
    it, bad = s.stats()
    print(f"   residual user {ru:.2e} item {ri:.2e}; dual_max {s.dual_max} short_max {s.short_max}; "
          f"CG iterations/row/epoch {it / (s.epochs_done + 1) / (train.shape[0] + train.shape[1]):.2f}, unconverged {bad}")
if world > 1:
    dist.destroy_process_group()
