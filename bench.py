#!/usr/bin/env python
"""bench.py -- the driver's benchmark contract for cymf_b200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

Metric (BASELINE.json): BPR triplet-updates/sec.  A "step" is one epoch of the BPR hot loop
(cymf/bpr.pyx:162-169) over the shuffled (user, positive) pairs of a synthetic ml-20m-shaped matrix
(138,493 x 26,744, 20 M nnz before the 10 % hold-out; BASELINE.json configs[2], K=128, SGD, f32 storage).
`value`   = applied (non-colliding) updates of all ranks / max-over-ranks device time, factors, pairs and
            CSR resident in HBM.  At N>1 BPR does not shard (every triplet may touch any
            row), so each rank trains an independent replica: "replicas only", scaling "weak".
`e2e`     = same metric through the reference-facing typed boundary `BPR._fit_bpr(users, positives, X, 1,
            lr, wd, ...)` with HOST numpy buffers: H2D of pairs, CSR and both f64 factors and D2H of the
            factors are inside the timed region every step.
`roofline`= algorithmic bytes of the epoch kernel (6*K*4+8 B per applied update) / its CUDA-event duration,
            against MEASURED_PEAKS.json's HBM copy bandwidth.
`cpu_baseline` = the compiled reference (oracle/_ref, OpenMP, all host cores) on a row-block sample.
`extra`   = the other hot-path kernels on their BASELINE.json configs (K=64 target run, Adam, ALS, GloVe, ...).
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

METRIC = "bpr_triplet_updates_per_sec"
UNIT = "updates/s"
WORKLOAD = "bpr-sgd-f32 K=128 on synthetic ml-20m-shaped CSR (138493x26744, 20M nnz, 10% held out)"
LR, WD, K_MAIN = 0.01, 0.01, 128
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


def time_als_device_pipeline(scale, K, steps, hbm):
    """BASELINE configs[4] pipeline at `scale` of 10 M x 1 M x 1 B nnz on ONE GPU: matrix generated on the device,
    X^T / row deal / blocks by prep.cu, device-initialised factors (tools/c5_als.py runs the full size, 1-8 GPUs)."""
    import torch
    from cymf_b200 import _lib
    from cymf_b200.synth import synth_implicit_device
    from cymf_b200.wmf import AlsSession
    U, I, nnz = int(10_000_000 * scale), int(1_000_000 * scale), int(1_000_000_000 * scale)
    ip, ix = synth_implicit_device(U, I, nnz, seed=104)
    real_nnz = int(ix.numel())
    g = torch.Generator(device="cuda")
    g.manual_seed(4321)
    ld = _lib.ld_for(K)
    W = (torch.rand((U, ld), device="cuda", generator=g) * 0.2 - 0.1) / K
    H = (torch.rand((I, ld), device="cuda", generator=g) * 0.2 - 0.1) / K
    W[:, K:] = 0
    H[:, K:] = 0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s = AlsSession((ip, ix, (U, I)), W, H, 0.01, 10.0, K=K, distributed=False)
    torch.cuda.synchronize()
    prep_s = time.perf_counter() - t0
    del W, H, ip, ix
    for _ in range(2):
        s.epoch()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        s.epoch()
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3 / steps
    algo = 2 * real_nnz * (K * 4 + 4) + (U + I) * (K * 4 + 8) + (U + I) * K * 4
    out = {"sec_per_epoch": sec, "shape": [U, I], "nnz": real_nnz, "K": K, "prepare_on_device_sec": prep_s,
           "algorithmic_GBps": algo / sec / 1e9, "frac_of_hbm_peak": algo / sec / 1e9 / hbm,
           "unconverged_rows": s.stats()[1],
           "full_size": "tools/c5_als.py: 0.90 s/epoch on 1 B200, 0.168 s on 8 (profiles/r1_c5_als_*.json)"}
    del s
    torch.cuda.empty_cache()
    return out


def time_bpr_e2e(train, users, positives, K, optimizer, steps, warmup, lr=LR, wd=WD):
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
    m = cymf.BPR(K, lr, optimizer, wd)
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
                   cg_max_iter=2 * K)
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
           "frac_of_hbm_peak": s.bytes_per_epoch / sec / 1e9 / (hbm * world)}
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
    accept = 1.0 - float(np.mean(np.diff(Xs.indptr))) / train.shape[1]   # E[1 - deg_u / I] over pairs ~ this
    return {"value": n * accept / per_epoch, "unit": UNIT, "cores": threads, "kind": "reference",
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
def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    train, users, positives = dataset("ml-20m")
    # each "step" = one epoch over the bounded sample; W warm-up epochs and K timed ones by differencing
    base = cpu_reference_bpr(train, users, positives, K_MAIN, "sgd", epochs=(args.warmup, args.warmup + args.steps))
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["sec_per_sample_epoch"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "reference computes in f64 on the host"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-extra", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--full", action="store_true", help="GloVe extra at the full 1e8 co-occurrences of configs[3]")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from cymf_b200 import _lib
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

    launches0 = _lib.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    sync_all()
    secs, applied, sess, per_step = time_bpr_device(train, users, positives, K_MAIN, "sgd", args.steps, args.warmup)
    sync_all()
    clocks = sampler.stop()
    launches = sess.timed_launches
    stats = torch.tensor([secs, float(applied)], dtype=torch.float64, device="cuda")
    if world > 1:
        tmax = stats.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        secs_all, applied_all = float(tmax[0]), float(stats[1])
    else:
        secs_all, applied_all = secs, float(applied)
    value = applied_all / secs_all
    bpu = sess.bytes_per_update
    kernel_s = float(np.mean(per_step))
    achieved = (applied / args.steps) * bpu / kernel_s / 1e9
    del sess
    torch.cuda.empty_cache()

    # end-to-end through the typed boundary with host buffers
    e_steps = max(2, min(args.steps, 5))
    sync_all()
    e_secs, e_applied, h2d, d2h = time_bpr_e2e(train, users, positives, K_MAIN, "sgd", e_steps, args.warmup)
    e = torch.tensor([e_secs, float(e_applied)], dtype=torch.float64, device="cuda")
    if world > 1:
        emax = e.clone()
        dist.all_reduce(emax, op=dist.ReduceOp.MAX)
        dist.all_reduce(e, op=dist.ReduceOp.SUM)
        e_value = float(e[1]) / float(emax[0])
    else:
        e_value = e_applied / e_secs

    extra = {}
    if rank == 0 and not args.no_extra:
        for tag, K, opt, dt in (("bpr_sgd_f32_k64", 64, "sgd", "float32"), ("bpr_adam_f32_k128", 128, "adam", "float32"),
                                ("bpr_sgd_f64_k128", 128, "sgd", "float64")):
            s_, a_, ss, ps = time_bpr_device(train, users, positives, K, opt, max(3, args.steps // 2), 3, dtype=dt)
            extra[tag] = {"updates_per_s": a_ / s_, "ms_per_epoch": 1e3 * s_ / len(ps),
                          "algorithmic_GBps": a_ * ss.bytes_per_update / s_ / 1e9,
                          "frac_of_hbm_peak": a_ * ss.bytes_per_update / s_ / 1e9 / hbm}
            del ss
            torch.cuda.empty_cache()
        extra["wmf_als_f32_k128_c5_scale0.1"] = time_als_device_pipeline(0.1, 128, 3, hbm)
        extra["evaluator_ml20m_k128"] = time_evaluator("ml-20m", 128, 3, hbm, with_reference=(world == 1 and not args.no_cpu))
        for tag, opt in (("relmf_sgd_f32_k128", "sgd"), ("relmf_adam_f32_k128", "adam")):
            extra[tag] = time_relmf_device(train, 128, opt, max(3, args.steps // 2), 3, hbm, 50_000_000)
            torch.cuda.empty_cache()
    if not args.no_extra:
        # WMF ALS shards over the ranks (row blocks, all-gather + Gram all-reduce): every rank takes part
        for tag, name, K in (("wmf_als_f32_k64_ml1m", "ml-1m", 64), ("wmf_als_f32_k128_ml20m", "ml-20m", 128)):
            if rank == 0:
                dataset(name)
            sync_all()
            res = time_als(dataset(name)[0], K, "float32", max(3, args.steps // 2), 3, sync_all, world, hbm)
            if rank == 0:
                extra[tag] = res
        if rank == 0:
            extra["glove_adagrad_f32_k300"] = time_glove(400_000, 100_000_000 if args.full else 20_000_000, 300,
                                                         max(3, args.steps // 2), 3, hbm)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            cpu = cpu_reference_bpr(train, users, positives, K_MAIN, "sgd")
        except Exception as ex:                                      # the checker is optional for the GPU number
            cpu = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {ex}"}

    if cpu is not None and not args.no_extra:
        extra["cpu_reference"] = cpu_reference_extras()
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": 1e3 * secs_all / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "optimizer": "sgd", "learning_rate": LR, "weight_decay": WD,
                           "pairs_per_epoch": int(users.shape[0]), "acceptance": applied / (args.steps * users.shape[0]),
                           "multi_gpu": "replicas only (BPR does not shard)" if world > 1 else "single GPU",
                           "l2": "inputs larger than L2: 144 MB of pairs are streamed per step next to 85 MB of "
                                 "factors (L2 = 126 MB); no explicit flush"},
                "clocks": clocks,
                "e2e": {"value": e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "steps": e_steps, "call": "cymf_b200.BPR._fit_bpr(users, positives, X, 1, lr, wd, 1, False)"},
                "gpu_launches": int(launches),
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                             # dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full capture of this
                             # command (profiles/r1_bpr_hogwild_k128_ncu_full.txt): 9.274 GB + 3.882 GB
                             "traffic": 13.156e9, "traffic_unit": "bytes/launch",
                             "algorithmic_bytes_per_launch": (applied / args.steps) * bpu,
                             "peak_source": peak_src, "kernel": "bpr_hogwild_kernel<float,SGD,32,1,RED>",
                             "bytes_per_update": bpu,
                             "note": "factors (85 MB) fit the 126 MB L2, so algorithmic bytes/s may exceed DRAM bytes/s"},
                "cpu_baseline": cpu, "extra": extra}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
