"""Worker for the multi-rank WMF tests (launched by torch.distributed.run).

  mode "gloo": CPU emulation of AlsSession's sharded half sweep -- same `_deal` / `_relabel` partition, same
               collective choreography (all-reduce of the K x K Gram partial, all-gather of the solved block)
               over the gloo backend, with the CPU oracle as the per-block solver.  Checks the host-side logic.
  mode "nccl": the real thing on GPUs: sharded cymf_b200.WMF vs a single-GPU run on rank 0.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def gloo_main():
    from scipy import sparse
    from cymf_b200.synth import synth_implicit
    from cymf_b200.wmf import _deal, _relabel
    from oracle import oracle
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    K, wd, weight = 12, 0.01, 10.0
    X = synth_implicit(97, 61, 1500, seed=3).tolil()
    X[5, :] = 0
    X = X.tocsr()
    X.eliminate_zeros()
    XT = X.T.tocsr()
    U, I = X.shape
    W0, H0 = oracle.init_factors(U, I, K)
    Wref, Href = oracle.wmf_fit(X, K, wd, weight, 2, W0.copy(), H0.copy())

    slot_u, Ru = _deal(np.diff(X.indptr), world)
    slot_i, Ri = _deal(np.diff(XT.indptr), world)
    assert sorted(slot_u[slot_u >= 0]) == list(range(U)) and sorted(slot_i[slot_i >= 0]) == list(range(I))
    new_u = np.empty(U, np.int64); new_u[slot_u[slot_u >= 0]] = np.flatnonzero(slot_u >= 0)
    new_i = np.empty(I, np.int64); new_i[slot_i[slot_i >= 0]] = np.flatnonzero(slot_i >= 0)
    blk_u = _relabel(X, slot_u[rank * Ru:(rank + 1) * Ru], new_i, slot_i.shape[0])
    blk_i = _relabel(XT, slot_i[rank * Ri:(rank + 1) * Ri], new_u, slot_u.shape[0])
    # nnz balance of the deal: no rank more than 15 % above the mean
    nnz = torch.tensor([float(blk_u.nnz), float(blk_i.nnz)])
    tot = nnz.clone(); dist.all_reduce(tot)
    assert (nnz <= 1.15 * tot / world + 8).all(), (nnz, tot)

    def dealt(M, slots):
        out = np.zeros((slots.shape[0], K)); out[slots >= 0] = M[slots[slots >= 0]]; return out
    W, H = dealt(W0, slot_u), dealt(H0, slot_i)

    def half(Xf, R, blk, Yf, Ry):
        yb = Yf[rank * Ry:(rank + 1) * Ry]
        g = torch.from_numpy(yb.T @ yb)                       # Gram partial of the own block ...
        dist.all_reduce(g)                                    # ... all-reduced (then + wd I inside the solver)
        assert np.allclose(g.numpy(), Yf.T @ Yf, rtol=1e-12, atol=1e-14)
        xb = Xf[rank * R:(rank + 1) * R].copy()
        oracle.als_half(blk.indptr, blk.indices, xb, Yf, wd, weight)
        out = torch.empty(Xf.shape, dtype=torch.float64)
        dist.all_gather_into_tensor(out, torch.from_numpy(xb))
        Xf[...] = out.numpy()

    for _ in range(2):
        half(W, Ru, blk_u, H, Ri)
        half(H, Ri, blk_i, W, Ru)
    Wg = np.empty_like(W0); Wg[slot_u[slot_u >= 0]] = W[slot_u >= 0]
    Hg = np.empty_like(H0); Hg[slot_i[slot_i >= 0]] = H[slot_i >= 0]
    assert np.abs(Wg - Wref).max() <= 1e-10 * np.abs(Wref).max()
    assert np.abs(Hg - Href).max() <= 1e-10 * np.abs(Href).max()
    assert not Wg[5].any()
    dist.destroy_process_group()
    print(f"rank {rank}: gloo sharded ALS == oracle")


def nccl_main():
    import cymf_b200 as cymf
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank()
    train, _ = cymf.synth.movielens_like("ml-1m")
    K = 64
    for dtype, tol in (("float64", 1e-9), ("float32", 2e-5)):
        m = cymf.WMF(K, 0.01, 10.0, dtype=dtype, distributed=True)
        m.fit(train, 2, 1, verbose=False)                      # sharded over all ranks
        s = cymf.WMF(K, 0.01, 10.0, dtype=dtype, distributed=False)
        s.fit(train, 2, 1, verbose=False)                      # every rank alone
        ew = np.abs(m.W - s.W).max() / np.abs(s.W).max()
        eh = np.abs(m.H - s.H).max() / np.abs(s.H).max()
        print(f"rank {rank} {dtype}: sharded ({m.gather_mode_}) vs single-GPU rel err W {ew:.2e} H {eh:.2e}", flush=True)
        assert ew <= tol and eh <= tol
        t = torch.from_numpy(m.W).cuda()
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), "ranks disagree on W"      # replicas are bit-identical across ranks
        n = cymf.WMF(K, 0.01, 10.0, dtype=dtype, peer_gather=False, distributed=True)
        n.fit(train, 2, 1, verbose=False)                      # same sharding, blocks exchanged by NCCL all-gather
        assert n.gather_mode_ == "nccl"
        # same factors whichever transport carries the exchange (the Gram partials are summed in rank order by the
        # peer-load kernel and in NCCL's order by all_reduce: last-bit differences in G only)
        for a_, b_ in ((n.W, m.W), (n.H, m.H)):
            assert np.abs(a_ - b_).max() <= (1e-11 if dtype == "float64" else 2e-6) * np.abs(b_).max(), "peer-store and NCCL paths differ"
    dist.destroy_process_group()


if __name__ == "__main__":
    {"gloo": gloo_main, "nccl": nccl_main}[sys.argv[1]]()
