"""`cymf.evaluator` on a B200: `Evaluator`, `AverageOverAllEvaluator` (`AoaEvaluator`) with the reference's
constructor and `evaluate(W, H, seed=1234) -> {"DCG@5": ..}` contract (cymf/evaluator.pyx:35-149).

What runs where
  host  : candidate construction.  One sequential mt19937 for the whole call with per-user rejection counts
          (evaluator.pyx:82,95-111) cannot be parallelised without changing the lists, so the C ABI generates
          them on the host (cymf_eval_candidates_host) -- ONCE per (evaluator, seed): `fit()` calls
          `evaluate(W, H)` every epoch with the default seed, and the lists do not depend on W or H.
  device: scores of all candidates (f64, k-ascending dot), exact ranks, DCG / Recall / MAP @k of every user
          (cymf_eval_rank_dev).  The final `.mean()` over ALL users (evaluator.pyx:135-137) is NumPy, as in
          the reference.

Tie order: the reference uses NumPy's unstable argsort and reverses it, which leaves the order of equal
scores undefined; here equal scores rank by descending candidate position (= reversed stable argsort).
`unbiased=True` (the IPS variant, whose propensity lookup indexes by candidate position, evaluator.pyx:116)
is outside this build's scope and raises NotImplementedError.
"""
import ctypes as C
import math

import numpy as np
from scipy import sparse

from . import _lib

__all__ = ["Evaluator", "AverageOverAllEvaluator", "AoaEvaluator", "UnbiasedEvaluator"]


class Evaluator(object):
    def __init__(self, X, X_train=None, metrics=["DCG", "Recall", "MAP"], k=5, num_negatives=100, unbiased=False):
        self.X = sparse.csr_matrix(X)
        self.user_positives = self.X.copy()
        if X_train is not None:
            self.user_positives += sparse.csr_matrix(X_train)
        self.X = self.X.astype(np.float64)
        self.user_positives = self.user_positives.astype(np.float64).tocsr()
        self.user_positives.sort_indices()
        self.propensity_scores = np.maximum(np.array(sparse.csr_matrix(X).mean(axis=0)).flatten(), 1e-4)
        self.metrics = metrics
        self.k = k
        self.num_negatives = int(num_negatives)
        self.unbiased = unbiased
        for m in self.metrics:
            if m not in ("DCG", "Recall", "MAP"):
                raise ValueError(f"unknown metric {m}")
        self._cand = {}          # seed -> device-resident candidate lists
        self._dev_csr = None
        self.last_order_ = None

    # ---- host: candidate lists (cached per seed) -----------------------------------------------------------
    def candidates(self, seed=1234):
        """(cand_ptr int64[U+1], cand_items int32[...]) exactly as evaluator.pyx:95-111 builds them."""
        U, I = self.X.shape
        tip = np.ascontiguousarray(self.X.indptr, np.int32)
        tix = np.ascontiguousarray(self.X.indices, np.int32)
        aip = np.ascontiguousarray(self.user_positives.indptr, np.int32)
        aix = np.ascontiguousarray(self.user_positives.indices, np.int32)
        n_eval = int((np.diff(tip) > 0).sum())
        cap = int(tix.shape[0] + n_eval * self.num_negatives)
        cand_ptr = np.empty(U + 1, np.int64)
        cand_items = np.empty(max(cap, 1), np.int32)
        p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        _lib.check(_lib.lib().cymf_eval_candidates_host(U, I, p(tip), p(tix), p(aip), p(aix), self.num_negatives,
                                                        int(seed), p(cand_ptr), p(cand_items), cap))
        return cand_ptr, cand_items[:cap]

    def _device_candidates(self, seed, dev):
        import torch
        # the lists depend on the seed, on num_negatives and on the two CSR structures: all of them are public
        # attributes, so all of them are part of the cache key (changing one after a first evaluate() rebuilds)
        key = (int(seed), str(dev), int(self.num_negatives), id(self.X), int(self.X.nnz), id(self.user_positives),
               int(self.user_positives.nnz))
        if key not in self._cand:
            cand_ptr, cand_items = self.candidates(seed)
            self._cand = {key: (torch.from_numpy(cand_ptr).to(dev), torch.from_numpy(cand_items).to(dev),
                                int(np.diff(cand_ptr).max()) if cand_ptr.shape[0] > 1 else 0)}
        csr_key = (str(dev), id(self.X), int(self.X.nnz))
        if self._dev_csr is None or self._dev_csr[0] != csr_key:
            self._dev_csr = (csr_key, torch.from_numpy(np.ascontiguousarray(self.X.indptr, np.int32)).to(dev))
        return self._cand[key], self._dev_csr[1]

    # ---- evaluate ------------------------------------------------------------------------------------------------
    def evaluate(self, W, H, seed=1234, return_order=False):
        """`Evaluator.evaluate(W, H, seed)` (evaluator.pyx:57-139).  Ranking is exact and deterministic: candidates are
        ordered by decreasing f64 score, and EQUAL scores rank by decreasing candidate position (the order the
        reference's `argsort()[::-1]` gives for runs of equal keys under a stable sort; NumPy's default quicksort is
        not stable, so on exact ties -- e.g. a user whose W row is all zero, as WMF leaves users without training
        entries -- the reference's own order is unspecified and may differ)."""
        if self.unbiased:
            raise NotImplementedError("UnbiasedEvaluator (IPS metrics, evaluator.pyx:115-123) is out of scope of "
                                      "cymf_b200; use AverageOverAllEvaluator")
        torch = _lib.require_cuda()
        dev = W.device if hasattr(W, "device") and not isinstance(W, np.ndarray) else torch.device("cuda")
        if type(self.k) == int:                                              # evaluator.pyx:84-85
            self.k = [self.k]
        ks = [int(k) for k in self.k]
        kmax = max(ks)
        U, I = self.X.shape

        def to_dev(a):
            if isinstance(a, np.ndarray):
                return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)   # evaluator.pyx:58-59
            return a.to(dev, torch.float64).contiguous()

        dW, dH = to_dev(W), to_dev(H)
        K = dW.shape[1]
        (cand_ptr, cand_items, max_c), tip = self._device_candidates(seed, dev)
        with torch.cuda.device(dev):
            d_ks = torch.tensor(ks, dtype=torch.int32, device=dev)
            d_lg = torch.tensor([math.log2(i + 1.0) for i in range(kmax)], dtype=torch.float64, device=dev)
            per_user = torch.empty((U, len(ks), 3), dtype=torch.float64, device=dev)
            order = torch.empty_like(cand_items) if return_order else None
            _lib.check(_lib.lib().cymf_eval_rank_dev(_lib.ptr(dW), _lib.ptr(dH), U, K, _lib.ptr(tip), _lib.ptr(cand_ptr),
                                                     _lib.ptr(cand_items), max_c, _lib.ptr(d_ks), len(ks),
                                                     _lib.ptr(d_lg), kmax, _lib.ptr(per_user), _lib.ptr(order),
                                                     _lib.stream_ptr()))
            buff = per_user.cpu().numpy()
        col = {"DCG": 0, "Recall": 1, "MAP": 2}
        out = {}
        for q, k in enumerate(ks):
            for metric in self.metrics:
                out[f"{metric}@{k}"] = buff[:, q, col[metric]].mean()          # evaluator.pyx:135-137
        if return_order:
            self.last_order_ = order.cpu().numpy()
        return out


class AverageOverAllEvaluator(Evaluator):
    def __init__(self, X, X_train=None, metrics=["DCG", "Recall", "MAP"], k=5, num_negatives=100):
        super(AverageOverAllEvaluator, self).__init__(X, X_train, metrics, k, num_negatives, unbiased=False)


AoaEvaluator = AverageOverAllEvaluator


class UnbiasedEvaluator(Evaluator):
    def __init__(self, X, X_train=None, metrics=["DCG", "Recall", "MAP"], k=5, num_negatives=100):
        super(UnbiasedEvaluator, self).__init__(X, X_train, metrics, k, num_negatives, unbiased=True)
