#!/usr/bin/env python
"""bench.py -- the driver's benchmark contract for cymf_b200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-extra] [--no-cpu]

Metric (BASELINE.json: "WMF ALS sec/epoch at 1/2/4/8 B200", reported as its reciprocal so that higher is better and
the driver's strong-scaling efficiency is value_N / (N value_1)): `wmf_als_epochs_per_sec`.  A "step" is one epoch
of `WMF._fit_als` (cymf/wmf.pyx:110-112: user half sweep + item half sweep of `_als`, wmf.pyx:136-174) at K=128 on
the synthetic ml-20m-shaped matrix (138,493 x 26,744, 20 M nnz before the 10 % hold-out) -- the configuration the
north star's "< 50 ms on 8 B200" target is quoted on.  At N>1 the SAME matrix is sharded by user / item row blocks
over the ranks: "scaling": "strong".
`value`   = epochs / max-over-ranks device time, CSR blocks and factors resident in HBM.
`e2e`     = the same metric through the public call `WMF.fit(X, 1)` with HOST (pinned) numpy buffers: upload of the
            CSR and both f64 factor matrices, device-side preparation, one epoch, download of both factors, every step.
`roofline`= algorithmic bytes of the row-solver launches (N (K s + 4) + rows (K s + 8) per half sweep, SURVEY.md
            8(d)) / their CUDA-event durations, against MEASURED_PEAKS.json's HBM copy bandwidth.
`parity`  = float64 relative residual of the reference's per-row normal equations (wmf.pyx:161-168) on sampled rows
            + the heaviest row of every rank's block, after one more user and item half sweep; max over ranks.
`cpu_baseline` = the compiled reference (oracle/_ref, OpenMP, all host cores) `WMF._als` on every 20th user row and
            every 20th item row, extrapolated by the row ratio (labelled).
`extra`   = the other hot-path kernels on their BASELINE.json configs: BPR (the other half of BASELINE.json's metric:
            K=128 / K=64 / Adam / f64 with their own roofline blocks and e2e), C1 epochs/s, WMF C2 and C5, GloVe
            C4, RelMF, evaluator.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "wmf_als_epochs_per_sec"
UNIT = "epochs/s"
WORKLOAD = ("wmf-als-f32 K=128 weight_decay=0.01 weight=10 on synthetic ml-20m-shaped CSR (138493x26744, 20M nnz, "
            "10% held out); one step = one epoch (user + item half sweep); rows sharded over the ranks")
ALS_WD, ALS_WEIGHT = 0.01, 10.0
CPU_STRIDE = 20            # the CPU arms solve every 20th user row and every 20th item row of the same matrix
CPU_SAMPLE = (f"compiled reference cymf.WMF(128, 0.01, 10.0)._als (cymf/wmf.pyx:136) on every {CPU_STRIDE}th user row and "
              f"every {CPU_STRIDE}th item row of the workload with the full fixed side, all host threads; epoch time = "
              f"t_user_rows x {CPU_STRIDE} + t_item_rows x {CPU_STRIDE} (row-ratio extrapolation)")
BPR_UNIT = "updates/s"
# dram__bytes_read.sum + dram__bytes_write.sum of the row-solver launches of one half sweep, from the committed
# ncu --set full capture of the same workload (profiles/r2_als_rows_ws_dual_ml20m_ncu_full.txt: user half 0.185 GB over
# its three launches, item half 0.513 GB; the fixed side lives in L2 at this shape, so DRAM traffic is far BELOW the
# algorithmic bytes)
ROOFLINE_TRAFFIC = 0.349e9
LR, WD, K_MAIN = 0.01, 0.01, 128


def config_block(world):
    """Identical in both arms (the reference arm computes in f64 on the host and is sampled; both facts are stated)."""
    return {"workload": WORKLOAD, "K": K_MAIN, "weight_decay": ALS_WD, "weight": ALS_WEIGHT,
            "cpu_arm_sample": CPU_SAMPLE,
            "l2": "no explicit flush: one epoch streams 18.7 GB of gathered rows (user side 71 MB + item side 14 MB of "
                  "factors fit the 126 MB L2, which is part of the workload's nature at this shape)"}
CACHE = os.environ.get("CYMF_BENCH_CACHE", "/tmp/cymf_b200_cache")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def dataset(name, with_test=False):
    """(train CSR, users, positives) of a MovieLens shape -- or (train, held-out test) -- cached on local disk."""
    from scipy import sparse
    from sklearn import utils
    from cymf_b200 import synth
    os.makedirs(CACHE, exist_ok=True)
    f = os.path.join(CACHE, f"{name}_v2.npz")
    if os.path.exists(f):
        z = np.load(f)
        train = sparse.csr_matrix((np.ones(z["indices"].shape[0]), z["indices"], z["indptr"]), shape=tuple(z["shape"]))
        if with_test:
            test = sparse.csr_matrix((np.ones(z["test_indices"].shape[0]), z["test_indices"], z["test_indptr"]),
                                     shape=tuple(z["shape"]))
            return train, test
        return train, z["users"], z["positives"]
    train, test = synth.movielens_like(name)
    np.random.seed(4321)
    users, positives = utils.shuffle(*train.nonzero())          # bpr.pyx:104
    users, positives = users.astype(np.int32), positives.astype(np.int32)
    tmp = f + f".{os.getpid()}.tmp.npz"
    np.savez(tmp, indptr=train.indptr, indices=train.indices, shape=np.array(train.shape), users=users,
             positives=positives, test_indptr=test.indptr, test_indices=test.indices)
    os.replace(tmp, f)
    return (train, test) if with_test else (train, users, positives)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        busy = [v for v in sm if v > 0.5 * (mx[0] if mx else 1)] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": mx[0] if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
def time_bpr_device(train, users, positives, K, optimizer, steps, warmup, dtype="float32", lr=LR, wd=WD):
    """Device-resident epochs: returns (seconds for `steps` epochs, applied updates in them, session)."""
    import torch
    from cymf_b200.bpr import BprSession
    from cymf_b200.host import init_factors
    W, H = init_factors(train.shape[0], train.shape[1], K)
    s = BprSession(W, H, users, positives, train, optimizer, dtype=dtype)
    for _ in range(warmup):
        s.epoch(lr, wd)
    torch.cuda.synchronize()
    a0 = s.applied()
    from cymf_b200 import _lib
    l0 = _lib.launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    evs[0].record()
    for t in range(steps):
        s.epoch(lr, wd)
        evs[t + 1].record()
    torch.cuda.synchronize()
    s.timed_launches = _lib.launch_count() - l0               # kernels of this library inside the timed region
    per_step = [evs[t].elapsed_time(evs[t + 1]) * 1e-3 for t in range(steps)]
    return sum(per_step), s.applied() - a0, s, per_step


def time_relmf_device(train, K, optimizer, steps, warmup, hbm, samples, dtype="float32"):
    """RelMF epochs of `samples` uniformly drawn cells (the reference draws U*I per epoch; a bounded count here)."""
    import torch
    from cymf_b200.relmf import RelmfSession, item_propensities
    from cymf_b200.host import init_factors
    W, H = init_factors(train.shape[0], train.shape[1], K)
    s = RelmfSession(W, H, train, item_propensities(train.astype(np.float64)), optimizer, dtype=dtype,
                     samples_per_epoch=samples)
    for _ in range(warmup):
        s.epoch(0.01, 0.01, 0.1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        s.epoch(0.01, 0.01, 0.1)
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3 / steps
    gbps = samples * s.bytes_per_update / sec / 1e9
    return {"samples_per_s": samples / sec, "ms_per_epoch": 1e3 * sec, "samples_per_epoch": samples,
            "cells": int(train.shape[0]) * int(train.shape[1]), "K": K, "optimizer": optimizer,
            "algorithmic_GBps": gbps, "frac_of_hbm_peak": gbps / hbm}


def time_evaluator(name, K, steps, hbm, with_reference):
    """`AverageOverAllEvaluator(test, train, k=5).evaluate(W, H)` (cymf/evaluator.pyx:57-139) on device-resident
    factors: per-user candidate scoring, exact ranking, DCG/Recall/MAP@5.  Candidate lists (host mt19937 stream) are
    built once per seed and timed separately."""
    import torch
    import cymf_b200 as cymf
    train, test = dataset(name, with_test=True)
    U, I = train.shape
    ev = cymf.evaluator.AverageOverAllEvaluator(test, train, k=5)
    g = torch.Generator(device="cuda")
    g.manual_seed(1)
    W = torch.randn((U, K), dtype=torch.float64, device="cuda", generator=g)
    H = torch.randn((I, K), dtype=torch.float64, device="cuda", generator=g)
    t0 = time.perf_counter()
    res = ev.evaluate(W, H)                                   # builds + uploads the candidate lists
    torch.cuda.synchronize()
    first = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(steps):
        res = ev.evaluate(W, H)
    torch.cuda.synchronize()
    sec = (time.perf_counter() - t0) / steps
    cand_ptr, _ = ev.candidates()
    n_cand = int(cand_ptr[-1])
    algo = n_cand * (8 * K + 4) + 8 * U * K                   # gathered H rows + candidate ids + W
    out = {"sec_per_call": sec, "users_per_s": U / sec, "candidates": n_cand, "K": K, "shape": [U, I],
           "first_call_sec_incl_candidate_lists": first, "algorithmic_GBps": algo / sec / 1e9,
           "frac_of_hbm_peak": algo / sec / 1e9 / hbm, "DCG@5_random_factors": res["DCG@5"],
           "note": "wall clock around evaluate(): kernel + D2H of the per-user metrics + NumPy mean"}
    if with_reference:
        try:
            sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
            import cymf as ref
            rows = 4000
            rev = ref.evaluator.AverageOverAllEvaluator(test[:rows], train[:rows], k=5)
            Wc, Hc = W[:rows].cpu().numpy(), H.cpu().numpy()
            t0 = time.perf_counter()
            rev.evaluate(Wc, Hc)
            out["reference_users_per_s"] = rows / (time.perf_counter() - t0)
            out["reference_sample"] = f"compiled reference Evaluator.evaluate on the first {rows} users"
        except Exception as ex:                              # noqa: BLE001
            out["reference_users_per_s"] = None
            out["reference_sample"] = f"failed: {ex}"
    return out


def time_als_device_pipeline(scale, K, steps, hbm, world=1, sync_all=None):
    """BASELINE configs[4] pipeline at `scale` of 10 M x 1 M x 1 B nnz: matrix generated on every rank's GPU (same
    seed, identical), X^T / row deal / this rank's blocks by prep.cu, device-initialised factors, row blocks sharded
    over the ranks; normal-equation residual of one more epoch as the parity scalar."""
    import torch
    import torch.distributed as dist
    from cymf_b200.synth import synth_implicit_device
    from cymf_b200.wmf import AlsSession
    sync_all = sync_all or torch.cuda.synchronize
    U, I, nnz = int(10_000_000 * scale), int(1_000_000 * scale), int(1_000_000_000 * scale)
    ip, ix = synth_implicit_device(U, I, nnz, seed=104)
    real_nnz = int(ix.numel())
    g = torch.Generator(device="cuda")
    g.manual_seed(4321)
    ld = (K + 31) // 32 * 32
    W = (torch.rand((U, ld), device="cuda", generator=g) * 0.2 - 0.1) / K
    H = (torch.rand((I, ld), device="cuda", generator=g) * 0.2 - 0.1) / K
    W[:, K:] = 0
    H[:, K:] = 0
    sync_all()
    t0 = time.perf_counter()
    s = AlsSession((ip, ix, (U, I)), W, H, 0.01, 10.0, K=K, distributed=world > 1)
    sync_all()
    prep_s = time.perf_counter() - t0
    del W, H, ip, ix
    torch.cuda.empty_cache()
    for _ in range(2):
        s.epoch()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        s.epoch()
    e1.record()
    sync_all()
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3 / steps], dtype=torch.float64, device="cuda")
    s.user_half()
    ru = s.residual("user", 12, seed=11 + s.rank)
    s.item_half()
    ri = s.residual("item", 12, seed=12 + s.rank)
    r = torch.tensor([ru, ri], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(r, op=dist.ReduceOp.MAX)
    sec = float(t[0])
    algo = 2 * real_nnz * (K * 4 + 4) + (U + I) * (K * 4 + 8) + (U + I) * K * 4
    out = {"sec_per_epoch": sec, "n_gpus": world, "shape": [U, I], "nnz": real_nnz, "K": K,
           "prepare_on_device_sec": prep_s, "algorithmic_GBps_all_gpus": algo / sec / 1e9,
           "frac_of_hbm_peak_per_gpu": algo / sec / 1e9 / (hbm * world), "unconverged_rows": s.stats()[1],
           "row_solver": s.row_solver, "gather": "peer-store" if s.peer else ("nccl" if s.dist else "single"),
           "parity_rel_residual": {"user_rows": float(r[0]), "item_rows": float(r[1]), "rows_checked_per_side_per_rank": 13},
           "hbm_peak_allocated_GB": torch.cuda.max_memory_allocated() / 1e9}
    del s
    torch.cuda.empty_cache()
    return out


def time_bpr_e2e(train, users, positives, K, optimizer, steps, warmup, lr=LR, wd=WD, dtype="float32"):
    """Through the typed boundary with host buffers: every step uploads inputs and downloads the factors."""
    import torch
    import cymf_b200 as cymf
    from cymf_b200.host import init_factors
    from scipy import sparse

    def pinned(a):                       # the step's host inputs live in pinned memory (bench contract)
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()

    users, positives = pinned(users), pinned(positives)
    train = sparse.csr_matrix((train.data, pinned(train.indices), pinned(train.indptr)), shape=train.shape, copy=False)
    train.has_sorted_indices = True
    m = cymf.BPR(K, lr, optimizer, wd, dtype=dtype)
    m.W, m.H = (pinned(a) for a in init_factors(train.shape[0], train.shape[1], K))
    m.valid_evaluator, m.early_stopping = None, False
    applied, h2d, d2h = 0, 0, 0
    for _ in range(max(1, min(warmup, 2))):
        m._fit_bpr(users, positives, train, 1, lr, wd, 1, False)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        m._fit_bpr(users, positives, train, 1, lr, wd, 1, False)
        applied += m.n_applied_
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    N, U = users.shape[0], train.shape[0]
    h2d = m.W.nbytes + m.H.nbytes + 8 * N + 8 * (U + 1) + 4 * train.indices.shape[0]
    d2h = m.W.nbytes + m.H.nbytes + 8
    return dt, applied, h2d, d2h


def time_als(train, K, dtype, steps, warmup, sync_all, world, hbm):
    """WMF ALS epochs (user + item half sweep), factors and CSR blocks resident; sharded when world > 1."""
    import torch
    import torch.distributed as dist
    from cymf_b200.wmf import AlsSession
    from cymf_b200.host import init_factors
    W, H = init_factors(train.shape[0], train.shape[1], K)
    s = AlsSession(train, W, H, 0.01, 10.0, dtype=dtype, cg_tol=1e-6 if dtype == "float32" else 1e-10,
                   cg_max_iter=2 * K, distributed=world > 1)
    for _ in range(warmup):
        s.epoch()
    sync_all()
    it0 = s.stats()[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        s.epoch()
    e1.record()
    sync_all()
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3 / steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t[0])
    rows = (train.shape[0] + train.shape[1]) / world
    out = {"sec_per_epoch": sec, "n_gpus": world, "nnz": int(train.nnz), "K": K, "dtype": dtype, "solver": s.solver,
           "gather": "peer-store (fused into the GEMM epilogue)" if s.peer else ("nccl all-gather" if s.dist else "single"),
           "cg_iterations_per_row": (s.stats()[0] - it0) / (steps * 2 * rows) * 2, "cg_tol": s.cg_tol,
           "unconverged_rows": s.stats()[1], "algorithmic_GBps": s.bytes_per_epoch / sec / 1e9,
           "frac_of_hbm_peak": s.bytes_per_epoch / sec / 1e9 / (hbm * world), "row_solver": s.row_solver}
    s.user_half()
    r = torch.tensor([s.residual("user", 12, seed=7 + s.rank)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(r, op=dist.ReduceOp.MAX)
    out["parity_rel_residual_user_rows"] = float(r[0])
    del s
    torch.cuda.empty_cache()
    return out


def synth_cooc_device(V, nnz, seed):
    """Device-side twin of cymf_b200.synth.synth_cooc (Zipf rows/cols, log-normal counts) for the 1e8-sample config."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    cdf = torch.cumsum(1.0 / (torch.arange(V, device="cuda", dtype=torch.float64) + 1.0), 0)
    cdf /= cdf[-1].clone()
    keys = torch.empty(0, dtype=torch.int64, device="cuda")
    while keys.numel() < nnz:
        m = int((nnz - keys.numel()) * 1.3) + 1024
        r = torch.searchsorted(cdf, torch.rand(m, generator=g, device="cuda", dtype=torch.float64)).clamp_(max=V - 1)
        c = torch.searchsorted(cdf, torch.rand(m, generator=g, device="cuda", dtype=torch.float64)).clamp_(max=V - 1)
        keys = torch.unique(torch.cat([keys, r * V + c]))
    keys = keys[torch.randperm(keys.numel(), generator=g, device="cuda")[:nnz]]       # shuffled, as fit() does
    counts = torch.exp(torch.randn(nnz, generator=g, device="cuda", dtype=torch.float64) * 1.5).clamp_(min=0.1)
    return (keys // V).to(torch.int32), (keys % V).to(torch.int32), counts


def time_glove(V, nnz, K, steps, warmup, hbm):
    """GloVe AdaGrad epochs on a synthetic V-word vocabulary (BASELINE.json configs[3] shape), f32, device-resident."""
    import torch
    from cymf_b200.glove import GloveSession
    c, x, n = synth_cooc_device(V, nnz, 103)
    rng = np.random.default_rng(103)
    W, H = rng.uniform(-.5, .5, (V, K)) / K, rng.uniform(-.5, .5, (V, K)) / K
    bw, bh = rng.uniform(-.5, .5, V) / K, rng.uniform(-.5, .5, V) / K
    s = GloveSession(c.cpu().numpy(), x.cpu().numpy(), n.cpu().numpy(), W, bw, H, bh, dtype="float32")
    del c, x, n
    for _ in range(warmup):
        s.epoch(0.05, 10.0, 0.75)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        last = s.epoch(0.05, 10.0, 0.75, want_loss=False)
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3 / steps
    out = {"samples_per_s": nnz / sec, "sec_per_epoch": sec, "V": V, "nnz": nnz, "K": K,
           "algorithmic_GBps": nnz * s.bytes_per_sample / sec / 1e9,
           "frac_of_hbm_peak": nnz * s.bytes_per_sample / sec / 1e9 / hbm}
    del s
    torch.cuda.empty_cache()
    return out


def cpu_reference_bpr(train, users, positives, K, optimizer, budget_s=20.0, threads=None, epochs=None):
    """The compiled reference (oracle/_ref) on a row-block sample; setup removed by differencing two fits."""
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    import cymf as ref                                                  # the reference build, not cymf_b200
    threads = threads or os.cpu_count()
    rows = min(train.shape[0], 16_000)
    Xs = train[:rows]
    n = Xs.nnz

    def fit(e):
        m = ref.BPR(K, LR, optimizer, WD)
        t = time.perf_counter()
        m.fit(Xs, e, threads, verbose=False)
        return time.perf_counter() - t

    e1, e2 = (1, 3) if epochs is None else epochs
    t1 = fit(e1)
    t2 = fit(e2)
    per_epoch = max((t2 - t1) / (e2 - e1), 1e-9)
    deg = np.diff(Xs.indptr).astype(np.float64)
    accept = 1.0 - float((deg * deg).sum()) / (max(n, 1) * train.shape[1])   # E over PAIRS of 1 - deg_u / I
    return {"value": n * accept / per_epoch, "unit": BPR_UNIT, "cores": threads, "kind": "reference",
            "sample": f"first {rows} users of the workload ({n} pairs/epoch), cymf.BPR(K={K},{optimizer}).fit "
                      f"num_threads={threads}, (t({e2} ep)-t({e1} ep))/{e2 - e1}; attempts x expected acceptance "
                      f"{accept:.4f}",
            "sec_per_sample_epoch": per_epoch}


def cpu_reference_extras():
    """The compiled reference (oracle/_ref) timed on the host for the other paths, on bounded samples:
    WMF on the full C2 shape, GloVe K=300 on 1 M samples, RelMF K=128 on a 500-user block.  Setup is removed by
    differencing two fits where the API has no per-epoch hook."""
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    import cymf as ref
    from cymf_b200 import synth
    threads = os.cpu_count()
    out = {}

    def guarded(tag, fn):
        try:
            out[tag] = fn()
        except Exception as ex:                              # noqa: BLE001 - a baseline must never cost the GPU line
            out[tag] = {"failed": repr(ex)}

    def wmf():
        train, _, _ = dataset("ml-1m")
        t = []
        for e in (0, 3):
            m = ref.WMF(64, 0.01, 10.0)
            t0 = time.perf_counter()
            m.fit(train, e, threads, verbose=False)
            t.append(time.perf_counter() - t0)
        return {"sec_per_epoch": max(t[1] - t[0], 1e-9) / 3, "cores": threads,
                "sample": "cymf.WMF(64).fit on the full ml-1m shape (C2), (t(3 epochs) - t(0 epochs)) / 3"}

    def glove():
        V, n, K = 400_000, 1_000_000, 300
        rng = np.random.default_rng(103)
        c = rng.integers(0, V, n).astype(np.int32)
        x = rng.integers(0, V, n).astype(np.int32)
        cnt = np.maximum(np.exp(rng.normal(0.0, 1.5, n)), 0.1)
        W, H = rng.uniform(-.5, .5, (V, K)) / K, rng.uniform(-.5, .5, (V, K)) / K
        bw, bh = rng.uniform(-.5, .5, V) / K, rng.uniform(-.5, .5, V) / K
        g = ref.GloVe(K, 0.05, 0.75, 10.0)
        t = []
        for e in (0, 2):
            t0 = time.perf_counter()
            g._fit_glove(c, x, cnt, W, bw, H, bh, e, 0.05, 10.0, 0.75, threads, False)
            t.append(time.perf_counter() - t0)
        return {"samples_per_s": 2 * n / max(t[1] - t[0], 1e-9), "cores": threads,
                "sample": "cymf.GloVe(300)._fit_glove on 1 M uniform samples over a 400 k vocabulary, t(2 ep) - t(0 ep)"}

    def relmf():
        train, _, _ = dataset("ml-20m")
        rows = 500
        Xs = train[:rows]
        t = []
        for e in (0, 2):
            m = ref.RelMF(128, 0.1, 0.01, "sgd", 0.01)
            t0 = time.perf_counter()
            m.fit(Xs, e, threads)
            t.append(time.perf_counter() - t0)
        n = rows * train.shape[1]
        return {"samples_per_s": 2 * n / max(t[1] - t[0], 1e-9), "cores": threads,
                "sample": f"cymf.RelMF(128, sgd).fit on the first {rows} users (dense {rows} x {train.shape[1]}), "
                          f"{n} sampled cells per epoch, t(2 ep) - t(0 ep)"}

    guarded("wmf_als_k64_ml1m", wmf)
    guarded("glove_adagrad_k300", glove)
    guarded("relmf_sgd_k128", relmf)
    return out


# ---------------------------------------------------------------------------------------------------------------
def cpu_reference_als(train, K, steps, warmup, threads=None):
    """The compiled reference's `WMF._als` (oracle/_ref, cymf/wmf.pyx:136-174) on a strided row sample of BOTH half
    sweeps of the workload; every step solves the same sample.  Returns the cpu_baseline block (value in epochs/s,
    extrapolated by the row ratio) and the per-step sample seconds."""
    sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
    import cymf as ref                                                  # the reference build, not cymf_b200
    from cymf_b200.host import init_factors
    threads = threads or os.cpu_count()
    U, I = train.shape
    W, H = init_factors(U, I, K)
    XT = train.T.tocsr()
    ru, ri = np.arange(0, U, CPU_STRIDE), np.arange(0, I, CPU_STRIDE)
    Xu, Xi = train[ru].tocsr(), XT[ri].tocsr()
    Wb, Hb = np.ascontiguousarray(W[ru]), np.ascontiguousarray(H[ri])
    m = ref.WMF(K, ALS_WD, ALS_WEIGHT)
    ipu, ixu = Xu.indptr.astype(np.int32), Xu.indices.astype(np.int32)
    ipi, ixi = Xi.indptr.astype(np.int32), Xi.indices.astype(np.int32)
    per = []
    for t in range(warmup + steps):
        t0 = time.perf_counter()
        m._als(ipu, ixu, Wb, H, threads)
        t1 = time.perf_counter()
        m._als(ipi, ixi, Hb, W, threads)
        t2 = time.perf_counter()
        if t >= warmup:
            per.append(((t1 - t0) * U / ru.shape[0] + (t2 - t1) * I / ri.shape[0], t2 - t0))
    epoch = float(np.mean([p[0] for p in per]))
    return {"value": 1.0 / epoch, "unit": UNIT, "cores": threads, "kind": "reference", "sample": CPU_SAMPLE,
            "sec_per_epoch_extrapolated": epoch, "sec_per_sample_step": float(np.mean([p[1] for p in per])),
            "sample_rows": [int(ru.shape[0]), int(ri.shape[0])], "sample_nnz": [int(Xu.nnz), int(Xi.nnz)],
            "dtype": "f64"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    train, _, _ = dataset("ml-20m")
    base = cpu_reference_als(train, K_MAIN, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["sec_per_sample_step"] * 1e3,
            "sec_per_epoch": base["sec_per_epoch_extrapolated"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_block(args.gpus), "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def time_als_main(train, K, steps, warmup, sync_all, world, hbm):
    """The headline: device-resident epochs of the sharded ALS session, per-epoch CUDA events, the row-solver
    launches timed with their own events, and the normal-equation residual of one more epoch."""
    import torch
    import torch.distributed as dist
    from cymf_b200 import _lib
    from cymf_b200.wmf import AlsSession
    from cymf_b200.host import init_factors
    W, H = init_factors(train.shape[0], train.shape[1], K)
    s = AlsSession(train, W, H, ALS_WD, ALS_WEIGHT, dtype="float32", cg_tol=1e-6, cg_max_iter=2 * K,
                   distributed=world > 1)
    for _ in range(warmup):
        s.epoch()
    sync_all()
    it0 = s.stats()[0]
    l0 = _lib.launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    evs[0].record()
    for t in range(steps):
        s.epoch()                                            # from the second epoch on: one CUDA-graph replay per epoch
        evs[t + 1].record()
    sync_all()
    launches = _lib.launch_count() - l0
    iters = s.stats()[0] - it0
    per_step = [evs[t].elapsed_time(evs[t + 1]) * 1e-3 for t in range(steps)]
    graphed = s._graph is not None
    if graphed:                                              # replays do not pass through the C entry points:
        launches = steps * s.graph_launches                  # kernels of this library captured in the graph, per replay
    # the row-solver launches alone, eagerly launched with their own CUDA events on the launching stream
    s.kernel_events = []
    for _ in range(max(2, min(steps, 5))):
        s.epoch()
    sync_all()
    kev, s.kernel_events = s.kernel_events, None
    k_bytes = float(sum(b for b, _, _ in kev))
    k_sec = float(sum(e0.elapsed_time(e1) for _, e0, e1 in kev)) * 1e-3
    t = torch.tensor([sum(per_step), k_sec], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    secs_all = float(t[0])
    # parity scalar: one more epoch, each half sweep checked right after it ran (the other side is still the one it
    # was solved against)
    s.user_half()
    ru = s.residual("user", 24, seed=1000 + s.rank)
    s.item_half()
    ri = s.residual("item", 24, seed=2000 + s.rank)
    r = torch.tensor([ru, ri], dtype=torch.float64, device="cuda")
    fin = torch.tensor([float(torch.isfinite(s.dW).all() and torch.isfinite(s.dH).all())], device="cuda")
    if world > 1:
        dist.all_reduce(r, op=dist.ReduceOp.MAX)
        dist.all_reduce(fin, op=dist.ReduceOp.MIN)
    rows = train.shape[0] + train.shape[1]
    out = {"secs_all": secs_all, "per_step": per_step, "launches": launches,
           "kernel_bytes_per_launch": k_bytes / max(len(kev), 1), "kernel_sec_per_launch": float(t[1]) / max(len(kev), 1),
           "kernel_launches": len(kev), "bytes_per_epoch": s.bytes_per_epoch,
           "cg_iterations_per_row": iters * world / (steps * rows), "unconverged_rows": s.stats()[1],
           "gather": "peer-store (fused into the GEMM epilogue)" if s.peer else ("nccl all-gather" if s.dist else "single"),
           "row_solver": s.row_solver, "peer_error": s.peer_error, "cuda_graph": graphed, "graph_error": s.graph_error,
           "parity": {"rel_residual": float(r.max()), "user_rows": float(r[0]), "item_rows": float(r[1]),
                      "rows_checked_per_side_per_rank": 25, "factors_finite": bool(fin.item() > 0),
                      "what": "max over ranks of |A x - b| / |b| in float64 for the reference's per-row system "
                              "(cymf/wmf.pyx:161-168) on 24 random rows + the heaviest row of each rank's block, "
                              "user rows after a user half sweep, item rows after an item half sweep; bar 1e-4"}}
    del s
    torch.cuda.empty_cache()
    return out


def time_als_e2e(train, K, steps, warmup, sync_all, world):
    """`cymf_b200.WMF(K).fit(X, 1)` per step with pinned HOST buffers: CSR + both f64 factor matrices go up, the
    device prepares X^T / the row deal / the blocks, one epoch runs, both factor matrices come back."""
    import torch
    import torch.distributed as dist
    import cymf_b200 as cymf
    from cymf_b200.host import init_factors
    from scipy import sparse

    def pinned(a):
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()

    X = sparse.csr_matrix((pinned(train.data.astype(np.float64)), pinned(train.indices), pinned(train.indptr)),
                          shape=train.shape, copy=False)
    X.has_sorted_indices = True
    m = cymf.WMF(K, ALS_WD, ALS_WEIGHT, distributed=world > 1)
    m.W, m.H = (pinned(a) for a in init_factors(train.shape[0], train.shape[1], K))
    for _ in range(max(1, min(warmup, 2))):
        m.fit(X, 1, 1, verbose=False)
    sync_all()
    t0 = time.perf_counter()
    for _ in range(steps):
        m.fit(X, 1, 1, verbose=False)
    sync_all()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    h2d, d2h = m.transfer_bytes_
    return float(dt[0]), int(h2d), int(d2h)


def bpr_block(train, users, positives, K, opt, dt, steps, hbm, traffic=None):
    """One BPR configuration, device-resident, with its own roofline block."""
    import torch
    s_, a_, ss, ps = time_bpr_device(train, users, positives, K, opt, steps, 3, dtype=dt)
    bpu = ss.bytes_per_update
    gbps = a_ * bpu / s_ / 1e9
    out = {"updates_per_s": a_ / s_, "ms_per_epoch": 1e3 * s_ / len(ps), "K": K, "optimizer": opt, "dtype": dt,
           "acceptance": a_ / (len(ps) * users.shape[0]),
           "roofline": {"bound": "hbm", "achieved": gbps, "peak": hbm, "unit": "GB/s", "frac": gbps / hbm,
                        "bytes_per_update": bpu, "traffic": traffic}}
    del ss
    torch.cuda.empty_cache()
    return out


def time_c1():
    """BASELINE configs[0]: BPR K=20 lr=0.01 wd=0.01, 30 epochs on the ml-100k shape through the public `fit()`
    (wall clock, everything included), in epochs/s next to the reference README's 98.46 it/s (README.md:62-66)."""
    import cymf_b200 as cymf
    train, _ = cymf.synth.movielens_like("ml-100k")
    out = {}
    for opt in ("adam", "sgd"):
        m = cymf.BPR(20, 0.01, opt, 0.01)
        m.fit(train, 2, 8, verbose=False)                     # library / allocator warm-up
        t0 = time.perf_counter()
        m = cymf.BPR(20, 0.01, opt, 0.01)
        m.fit(train, 30, 8, verbose=False)
        dt = time.perf_counter() - t0
        out[opt] = {"epochs_per_s_incl_setup": 30 / dt, "sec_for_30_epochs": dt}
    out["reference_readme_it_per_s"] = 98.46
    out["note"] = "reference figure: real ml-100k, Adam, 8 threads, unstated hardware (README.md:66); context only"
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--als-only", action="store_true", help="skip the non-ALS extras (BPR, GloVe, RelMF, evaluator)")
    ap.add_argument("--glove-samples", type=int, default=100_000_000, help="co-occurrences of the GloVe extra (configs[3]: 1e8)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if rank == 0:
        dataset("ml-20m")                                   # one rank builds the cache, the others read it
    if world > 1:
        dist.barrier()
    train, users, positives = dataset("ml-20m")
    hbm, peak_src = peaks()

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    sync_all()
    main_res = time_als_main(train, K_MAIN, args.steps, args.warmup, sync_all, world, hbm)
    sync_all()
    clocks = sampler.stop()
    secs_all = main_res["secs_all"]
    value = args.steps / secs_all
    k_gbps = main_res["kernel_bytes_per_launch"] / max(main_res["kernel_sec_per_launch"], 1e-12) / 1e9

    e_steps = max(2, min(args.steps, 5))
    e_secs, h2d, d2h = time_als_e2e(train, K_MAIN, e_steps, args.warmup, sync_all, world)

    def guarded(tag, fn):
        try:
            extra[tag] = fn()
        except Exception as ex:                              # noqa: BLE001 - an extra must never cost the headline line
            extra[tag] = {"failed": repr(ex)}
        torch.cuda.empty_cache()

    extra = {}
    x_steps = max(3, args.steps // 2)
    if not args.no_extra:
        # sharded over the ranks like the headline: every rank takes part
        for tag, name, K in (("wmf_als_f32_k64_ml1m_c2", "ml-1m", 64),):
            if rank == 0:
                dataset(name)
            sync_all()
            res = time_als(dataset(name)[0], K, "float32", x_steps, 3, sync_all, world, hbm)
            if rank == 0:
                extra[tag] = res
        if world > 1:
            # BASELINE configs[4] at FULL size (10 M x 1 M x ~1 B nnz, generated on the device), sharded
            try:
                res = time_als_device_pipeline(1.0, 128, 2, hbm, world, sync_all)
            except Exception as ex:                          # noqa: BLE001
                res = {"failed": repr(ex)}
            if rank == 0:
                extra["wmf_als_f32_k128_c5_full"] = res
    if rank == 0 and not args.no_extra and not args.als_only:
        # BPR: the other half of BASELINE.json's metric (configs[2] and the K=64 target run), one GPU
        guarded("bpr_sgd_f32_k128_c3", lambda: bpr_block(train, users, positives, 128, "sgd", "float32", x_steps, hbm,
                                                        traffic={"bytes_per_launch": 13.156e9,
                                                                 "source": "profiles/r1_bpr_hogwild_k128_ncu_full.txt"}))
        guarded("bpr_sgd_f32_k64", lambda: bpr_block(train, users, positives, 64, "sgd", "float32", x_steps, hbm))
        guarded("bpr_adam_f32_k128", lambda: bpr_block(train, users, positives, 128, "adam", "float32", x_steps, hbm))
        guarded("bpr_sgd_f64_k128", lambda: bpr_block(train, users, positives, 128, "sgd", "float64", x_steps, hbm))
        guarded("bpr_sgd_f64_k64", lambda: bpr_block(train, users, positives, 64, "sgd", "float64", x_steps, hbm))

        def bpr_e2e(dtype):
            e_s, e_a, bh2d, bd2h = time_bpr_e2e(train, users, positives, K_MAIN, "sgd", 3, 2, dtype=dtype)
            return {"updates_per_s": e_a / e_s, "h2d_bytes_per_step": int(bh2d), "d2h_bytes_per_step": int(bd2h),
                    "call": f"cymf_b200.BPR(dtype={dtype!r})._fit_bpr(users, positives, X, 1, lr, wd, 1, False)"}
        guarded("bpr_e2e_f32_k128", lambda: bpr_e2e("float32"))
        guarded("bpr_e2e_f64_k128", lambda: bpr_e2e("float64"))
        guarded("bpr_c1_ml100k_k20_30_epochs", time_c1)
        if world == 1:
            guarded("wmf_als_f32_k128_c5_scale0.1", lambda: time_als_device_pipeline(0.1, 128, 3, hbm, 1, sync_all))
        guarded("evaluator_ml20m_k128", lambda: time_evaluator("ml-20m", 128, 3, hbm,
                                                               with_reference=(world == 1 and not args.no_cpu)))
        for tag, opt in (("relmf_sgd_f32_k128", "sgd"), ("relmf_adam_f32_k128", "adam")):
            guarded(tag, lambda opt=opt: time_relmf_device(train, 128, opt, x_steps, 3, hbm, 50_000_000))
        guarded("glove_adagrad_f32_k300_c4", lambda: time_glove(400_000, args.glove_samples, 300, x_steps, 3, hbm))

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cpu = cpu_reference_als(train, K_MAIN, 3, 1)
        except Exception as ex:                                      # the checker is optional for the GPU number
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {ex}"}
        if not args.no_extra:
            try:
                extra["cpu_reference"] = cpu_reference_extras()
                extra["cpu_reference"]["bpr_sgd_k128_c3"] = cpu_reference_bpr(train, users, positives, K_MAIN, "sgd")
            except Exception as ex:                                  # noqa: BLE001
                extra["cpu_reference"] = {"failed": repr(ex)}
    if rank == 0:
        frac = k_gbps / hbm
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * secs_all / args.steps, "sec_per_epoch": secs_all / args.steps,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config_block(world),
                "parity": main_res["parity"],
                "clocks": clocks,
                "e2e": {"value": e_steps / e_secs, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "steps": e_steps, "call": "cymf_b200.WMF(128, 0.01, 10.0).fit(X, 1, 1, verbose=False), warm-started"},
                "gpu_launches": int(main_res["launches"]),
                "roofline": {"bound": "hbm", "achieved": k_gbps, "peak": hbm * 1.0, "unit": "GB/s", "frac": frac,
                             "traffic": ROOFLINE_TRAFFIC, "traffic_unit": "bytes/launch",
                             "traffic_source": "profiles/r2_als_rows_ws_dual_ml20m_ncu_full.txt: dram__bytes of the row-solver launches of one half "
                                               "sweep (user: ws 0.149 + dual 0.020 + 0.016 GB, item: ws 0.513 GB), mean of the two",
                             "algorithmic_bytes_per_launch": main_res["kernel_bytes_per_launch"],
                             "sec_per_launch": main_res["kernel_sec_per_launch"],
                             "launches_timed": main_res["kernel_launches"],
                             "peak_source": peak_src,
                             "kernel": "tc::als_rows_ws_kernel<128> + tc::als_rows_dual_kernel<128,{64,32}> (the row solvers of a half sweep)",
                             "per_gpu": "max-over-ranks kernel time against one GPU's peak: each rank moves its own block",
                             "epoch_algorithmic_GBps_all_gpus": main_res["bytes_per_epoch"] * args.steps / secs_all / 1e9,
                             "epoch_frac_of_aggregate_peak": main_res["bytes_per_epoch"] * args.steps / secs_all / 1e9 / (hbm * world),
                             "note": "the fixed side (<= 71 MB) fits the 126 MB L2 at this shape, so algorithmic bytes/s "
                                     "may exceed DRAM bytes/s"},
                "als": {k: main_res[k] for k in ("cg_iterations_per_row", "unconverged_rows", "gather", "row_solver",
                                                 "peer_error", "cuda_graph", "graph_error")},
                "cpu_baseline": cpu, "extra": extra}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
