"""Dev tool: A/B of the tail threshold of a sharded half sweep (rows holding more than 1 / divisor of their block's
entries take the direct solve, cymf_als_heavy_rows_dev) in ONE process per rank:
    torchrun --nproc-per-node N tools/als_tail_ab.py [ml-20m] [128] [256,148,100]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import cymf_b200 as cymf
from cymf_b200.wmf import AlsSession
from cymf_b200.host import init_factors

name, K = (sys.argv[1], int(sys.argv[2])) if len(sys.argv) > 2 else ("ml-20m", 128)
divs = [int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "256,148").split(",")]
world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
train, _ = cymf.synth.movielens_like(name)
W, H = init_factors(train.shape[0], train.shape[1], K)
for div in divs:
    os.environ["CYMF_ALS_TAIL_DIVISOR"] = str(div)
    s = AlsSession(train, W, H, 0.01, 10.0, cg_tol=1e-6, cg_max_iter=2 * K, distributed=world > 1)
    for _ in range(4):
        s.epoch()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        s.epoch()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 10], device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    s.user_half(); ru = s.residual("user", 8, seed=1)
    s.item_half(); ri = s.residual("item", 8, seed=2)
    if s.rank == 0:
        nh = (s.heavy_u[0] if s.heavy_u else 0, s.heavy_i[0] if s.heavy_i else 0)
        print(f"{name} K={K} world={world} tail divisor {div}: {float(t):.3f} ms/epoch (max over ranks), heavy rows on rank 0 "
              f"(user, item) {nh}, residual {max(ru, ri):.1e}", flush=True)
    del s
if world > 1:
    dist.destroy_process_group()
