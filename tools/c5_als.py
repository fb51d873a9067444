"""BASELINE.json configs[4]: WMF ALS K=128 on a synthetic 10 M users x 1 M items x ~1 B nnz matrix, sharded by
user / item row blocks over the ranks (1, 2, 4 or 8 B200; `torchrun --nproc-per-node N tools/c5_als.py`).

Nothing of this size ever exists on the host: the matrix is generated on every rank's GPU in user blocks from a
seeded torch generator (identical on all ranks), X^T / the row deal / the relabelled row blocks are built by
cymf_b200/csrc/prep.cu, the factors are initialised on the device.  Timing: CUDA events around whole epochs, max
over ranks.  Parity at this size is checked through the property the reference's dgesv solve guarantees
(cymf/wmf.pyx:161-168): every sampled row satisfies its normal equations
    (Y^T Y + wd I + (w - 1) sum_{c in row} y_c y_c^T) x = w sum_{c in row} y_c
evaluated in float64 from the device-resident factors, relative residual <= 1e-4.

    python tools/c5_als.py [--scale 1.0] [--epochs 3] [--K 128]
`--scale s` shrinks users, items and nnz by s (s = 0.01 is a 100 k x 10 k x 10 M smoke configuration).
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from cymf_b200.synth import synth_implicit_device as synth_c5_device  # noqa: E402


def residuals(sess, side, sample, seed):
    """Checker for sizes no CPU oracle can reach: `AlsSession.residual` (relative f64 residual of the reference's
    per-row normal equations, cymf/wmf.pyx:161-168, on sampled rows + the heaviest row of the block)."""
    return sess.residual(side, sample, seed)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--K", type=int, default=128)
    ap.add_argument("--check", type=int, default=24, help="rows per side whose normal equations are verified")
    ap.add_argument("--stage-rows", type=int, default=0, help="item vectors staged in shared memory per CTA (0 = auto)")
    ap.add_argument("--heavy-min", type=int, default=4096, help="rows with at least this many entries are solved directly")
    ap.add_argument("--no-overlap", action="store_true", help="row-class kernels back to back on one stream")
    ap.add_argument("--out", default=None)
    ap.add_argument("--cg-tol", type=float, default=1e-6, help="1e10 = no CG iterations (times the per-row Gram build alone)")
    args = ap.parse_args()

    import torch.distributed as dist
    from cymf_b200 import _lib
    from cymf_b200.wmf import AlsSession
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    U, I, nnz, K = int(10_000_000 * args.scale), int(1_000_000 * args.scale), int(1_000_000_000 * args.scale), args.K
    dev = torch.device("cuda", local)

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    t0 = time.perf_counter()
    indptr, indices = synth_c5_device(U, I, nnz, seed=104)
    sync()
    t_gen = time.perf_counter() - t0
    real_nnz = int(indices.numel())

    g = torch.Generator(device=dev)
    g.manual_seed(4321)
    ld = (K + 31) // 32 * 32
    W = (torch.rand((U, ld), device=dev, generator=g) * 0.2 - 0.1) / K         # wmf.pyx:88-92 on the device, f32
    H = (torch.rand((I, ld), device=dev, generator=g) * 0.2 - 0.1) / K
    W[:, K:] = 0
    H[:, K:] = 0
    launches0 = _lib.launch_count()
    t0 = time.perf_counter()
    sess = AlsSession((indptr, indices, (U, I)), W, H, 0.01, 10.0, K=K, dtype="float32", cg_tol=args.cg_tol, cg_max_iter=2 * K,
                      stage_rows=args.stage_rows, overlap_classes=not args.no_overlap, heavy_min=args.heavy_min,
                      distributed=world > 1)
    sync()
    t_prep = time.perf_counter() - t0
    prep_launches = _lib.launch_count() - launches0
    del W, H, indptr, indices
    torch.cuda.empty_cache()

    for _ in range(args.warmup):
        sess.epoch()
    sync()
    it0 = sess.stats()[0]
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.epochs + 1)]
    evs[0].record()
    for e in range(args.epochs):
        sess.epoch()
        evs[e + 1].record()
    sync()
    per_epoch = torch.tensor([evs[e].elapsed_time(evs[e + 1]) * 1e-3 for e in range(args.epochs)],
                             dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(per_epoch, op=dist.ReduceOp.MAX)
    iters, unconverged = sess.stats()

    # parity property at full size: one more epoch, checking each half sweep right after it ran
    sess.user_half()
    res_u = residuals(sess, "user", args.check, seed=rank)
    sess.item_half()
    res_i = residuals(sess, "item", args.check, seed=rank + 100)
    res = torch.tensor([res_u, res_i], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(res, op=dist.ReduceOp.MAX)
    finite = bool(torch.isfinite(sess.dW).all() and torch.isfinite(sess.dH).all())

    if rank == 0:
        sec = float(per_epoch.mean())
        s = 4
        algo = 2 * real_nnz * (K * s + 4) + (U + I) * (K * s + 8) + (U + I) * K * s
        hbm = 6456.2
        try:
            hbm = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                    "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            pass
        line = {"workload": f"wmf-als-f32 K={K} on synthetic {U}x{I} ({real_nnz} nnz), row-block sharded",
                "n_gpus": world, "sec_per_epoch": sec, "sec_per_epoch_each": [float(x) for x in per_epoch.tolist()],
                "epochs_timed": args.epochs, "warmup": args.warmup,
                "cg_iterations_per_row": (iters - it0) / (args.epochs * (U + I)) * world, "unconverged_rows": unconverged,
                "algorithmic_bytes_per_epoch": algo, "algorithmic_GBps_all_gpus": algo / sec / 1e9,
                "frac_of_hbm_peak_per_gpu": algo / sec / 1e9 / hbm / world,
                "stage_rows": args.stage_rows, "heavy_min": args.heavy_min,
                "heavy_rows": [h[0] if h else 0 for h in (sess.heavy_u, sess.heavy_i)], "gather": "peer-store" if sess.peer else ("nccl" if sess.dist else "single"),
                "generate_s": t_gen, "prepare_s": t_prep, "prepare_kernel_launches": prep_launches,
                "normal_equation_rel_residual_max": {"user_rows": float(res[0]), "item_rows": float(res[1]),
                                                     "rows_checked_per_side_per_rank": args.check + 1},
                "factors_finite": finite,
                "hbm_peak_allocated_GB": torch.cuda.max_memory_allocated() / 1e9}
        print(json.dumps(line))
        if args.out:
            with open(args.out, "w") as f:
                f.write(json.dumps(line) + "\n")
        assert finite and (float(res.max()) <= 1e-4 or args.cg_tol > 1e-6), line
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
