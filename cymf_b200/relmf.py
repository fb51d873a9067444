"""`cymf.RelMF` on a B200: same constructor, `fit` signature, attributes and error behaviour as the reference
class (cymf/relmf.pyx:37-171); the prange loop of `_fit_relmf` (relmf.pyx:143-148) runs as CUDA kernels
(cymf_b200/csrc/relmf.cu) reached through the C ABI of include/cymf_b200.h.

Host logic kept in Python/NumPy as in the reference: input coercion (relmf.pyx:76-80), propensities
(relmf.pyx:90), seeded init (relmf.pyx:92-96), per-epoch validation and early stopping (relmf.pyx:154-171).
One deliberate difference: the reference densifies X (`X.toarray()`, 29.6 GB at the ml-20m shape) and reads
X[u, i] from the dense copy; here X stays CSR and the kernel looks the cell up in the user's sorted row --
absent cells read 0.0, stored cells their value, exactly what the dense copy holds.

Keyword-only extras (not in the reference): `mode` ("hogwild" | "replay"), `dtype`, `scatter`, `seed`,
`max_inflight`, `device` with the meaning they have for `cymf_b200.BPR`, and `samples_per_epoch` (default U*I,
relmf.pyx:121).
"""
import ctypes as C

import numpy as np
from scipy import sparse

from . import _lib
from .host import init_missing_factors, run_epochs


def item_propensities(X):
    """relmf.pyx:90 on a CSR matrix: max(mean_u X[u,i] / max_i mean_u X[u,i], 1e-5) ** 0.5.

    The reference reduces a dense array along axis 0, i.e. adds rows in ascending u; the stored entries of a
    canonical CSR visited in order give the same partial sums (absent cells add 0.0), so the result is identical."""
    U, I = X.shape
    col_sum = np.bincount(X.indices, weights=X.data, minlength=I)     # sequential adds in CSR (ascending-u) order
    mean = col_sum / U
    return np.maximum(mean / mean.max(), 1e-5) ** 0.5


class RelMF(object):
    """
    Relevance Matrix Factorization (Rel-MF), https://arxiv.org/pdf/1909.03601.pdf

    Attributes:
        num_components (int): A dimensionality of latent vector
        clip_value (double): lower clip of the propensity score
        learning_rate (double): A learning rate
        optimizer (str): 'adam', 'adagrad' or 'sgd'
        weight_decay (double): A coefficient of weight decay
        W (np.ndarray[double, ndim=2]): User latent vectors
        H (np.ndarray[double, ndim=2]): Item latent vectors
    """

    def __init__(self, num_components=20, clip_value=0.1, learning_rate=0.001, optimizer="adam", weight_decay=0.01, *,
                 mode="hogwild", dtype="float32", scatter="auto", seed=1234, max_inflight=None, device=None,
                 samples_per_epoch=None):
        self.num_components = int(num_components)
        self.clip_value = float(clip_value)
        self.learning_rate = float(learning_rate)
        self.optimizer = optimizer
        self.weight_decay = float(weight_decay)
        self.W = None
        self.H = None
        if self.optimizer not in ("sgd", "adagrad", "adam"):
            raise Exception(f"{self.optimizer} is invalid.")            # relmf.pyx:65-66
        if mode not in ("hogwild", "replay"):
            raise ValueError("mode must be 'hogwild' or 'replay'")
        if dtype not in _lib.DTYPES:
            raise ValueError("dtype must be 'float32' or 'float64'")
        if scatter not in ("auto", "store", "red"):
            raise ValueError("scatter must be 'auto', 'store' or 'red'")
        self.mode, self.dtype, self.scatter = mode, dtype, scatter
        self.seed = int(seed)
        self.max_inflight = max_inflight
        self.device = device
        self.samples_per_epoch = samples_per_epoch

    def fit(self, X, num_epochs=10, num_threads=1, valid_evaluator=None, early_stopping=False, verbose=False):
        """
        Training RelMF model with Gradient Descent.

        Args:
            X: A user-item interaction matrix (scipy sparse or dense ndarray).
            num_epochs (int): A number of epochs.
            num_threads (int): accepted for signature compatibility; the GPU grid replaces the thread pool.
            verbose (bool): Whether to show the progress of training.
        """
        if X is None:
            raise ValueError()
        X = sparse.csr_matrix(X).astype(np.float64)                    # relmf.pyx:76-80 without the dense copy
        X.sum_duplicates()
        X.sort_indices()

        self.valid_evaluator = valid_evaluator
        self.valid_dcg = -np.inf
        self.count = 0
        self.early_stopping = early_stopping

        propensities = item_propensities(X)                            # relmf.pyx:90
        init_missing_factors(self, X.shape[0], X.shape[1])             # relmf.pyx:92-96
        return self._fit_relmf(X, propensities, num_epochs, num_threads, verbose)

    def _fit_relmf(self, X, propensities, num_epochs, num_threads, verbose):
        """Device replacement of `RelMF._fit_relmf` (relmf.pyx:107-171); X is CSR (or anything csr_matrix accepts)."""
        self.W = np.ascontiguousarray(self.W, dtype=np.float64)
        self.H = np.ascontiguousarray(self.H, dtype=np.float64)
        if not hasattr(self, "valid_dcg"):
            self.valid_dcg = -np.inf
        sess = RelmfSession(self.W, self.H, X, propensities, self.optimizer, mode=self.mode, dtype=self.dtype,
                            scatter=self.scatter, seed=self.seed, max_inflight=self.max_inflight, device=self.device,
                            samples_per_epoch=self.samples_per_epoch)
        run_epochs(self, sess, num_epochs,
                   lambda: sess.epoch(self.learning_rate, self.weight_decay, self.clip_value), verbose, ncols=100)
        self.n_samples_ = sess.n * sess.epochs_done


class RelmfSession(object):
    """Device-resident state of one `_fit_relmf` call: factors (+ optimizer state, rebuilt per fit as in
    relmf.pyx:129-137), the CSR of X with its values, the propensities.  `epoch()` enqueues one pass of U*I
    sampled cells (relmf.pyx:143-148) on the current CUDA stream."""

    def __init__(self, W, H, X, propensities, optimizer, *, mode="hogwild", dtype="float32", scatter="auto",
                 seed=1234, max_inflight=None, device=None, samples_per_epoch=None):
        torch = _lib.require_cuda()
        self._L = _lib.lib()
        self.dev = dev = torch.device(device if device is not None else "cuda")
        X = sparse.csr_matrix(X)
        if not X.has_sorted_indices:
            X = X.sorted_indices()
        self.U, self.I = U, I = X.shape
        self.K = K = W.shape[1]
        self.n = int(samples_per_epoch) if samples_per_epoch is not None else U * I      # relmf.pyx:121
        self.replay = mode == "replay"
        self.dtype = _lib.F64 if self.replay else _lib.DTYPES[dtype]
        self.opt = opt = _lib.OPTIMIZERS[optimizer]
        tdt = torch.float64 if self.dtype == _lib.F64 else torch.float32
        self.ld = ld = _lib.ld_for(K)
        data = np.ascontiguousarray(X.data, np.float64)
        binary = bool((data == 1.0).all())                        # implicit feedback: no value array on the device
        with torch.cuda.device(dev):
            self.d_indptr = torch.from_numpy(X.indptr.astype(np.int64)).to(dev, non_blocking=True)
            self.d_indices = torch.from_numpy(np.ascontiguousarray(X.indices, np.int32)).to(dev, non_blocking=True)
            self.d_values = None if binary else torch.from_numpy(data).to(dev).to(tdt)
            self.d_prop = torch.from_numpy(np.ascontiguousarray(np.asarray(propensities).ravel(), np.float64)
                                           ).to(dev).to(tdt)
            self.dW = _lib.upload_factor(W, self.dtype, dev)
            self.dH = _lib.upload_factor(H, self.dtype, dev)
            self.state = []
            if opt == _lib.ADAGRAD:
                self.state = [torch.ones((U, ld), dtype=tdt, device=dev), torch.ones((I, ld), dtype=tdt, device=dev)]
            elif opt == _lib.ADAM:
                self.state = [torch.zeros((n, ld), dtype=tdt, device=dev) for n in (U, I, U, I)]
        sp = [_lib.ptr(t) for t in self.state] + [None] * (4 - len(self.state))
        self.f = _lib.Factors(_lib.ptr(self.dW), _lib.ptr(self.dH), sp[0], sp[1], sp[2], sp[3])
        # f32: 128-bit reductions cost the same as stores and keep every concurrent sample of a row (parameters
        # and optimizer state alike); f64 reductions are scalar, so f64 keeps plain stores unless asked.
        self.scatter = {"auto": 1 if self.dtype == _lib.F32 else 0, "store": 0, "red": 1}[scatter]
        # bounds Hogwild staleness: at most ~4 samples of any one row in flight on small matrices (Adam's first
        # moment, advanced by concurrent increments, needs fewer than ten); large matrices fill the machine
        self.inflight = (int(max_inflight) if max_inflight is not None
                         else max(256, min(self.n // 256, 4 * min(U, I))))
        self.seed = int(seed)
        self.gen = _lib.HostRng(1234) if self.replay else None    # relmf.pyx:127: one generator per fit
        self.epochs_done = 0

    def epoch(self, learning_rate, weight_decay, clip_value):
        import torch
        L, f = self._L, self.f
        with torch.cuda.device(self.dev):
            stream = _lib.stream_ptr()
            if self.replay:
                cells = torch.from_numpy(self.gen.below64(self.U * self.I, self.n)).to(self.dev)
                _lib.check(L.cymf_relmf_replay_epoch_dev(
                    C.byref(f), self.opt, _lib.ptr(cells), self.n, _lib.ptr(self.d_indptr), _lib.ptr(self.d_indices),
                    _lib.ptr(self.d_values), _lib.ptr(self.d_prop), self.U, self.I, self.K, self.ld,
                    learning_rate, weight_decay, clip_value, stream))
                self._keep = cells
            else:
                _lib.check(L.cymf_relmf_hogwild_epoch_dev(
                    C.byref(f), self.dtype, self.opt, self.scatter, _lib.ptr(self.d_indptr), _lib.ptr(self.d_indices),
                    _lib.ptr(self.d_values), _lib.ptr(self.d_prop), self.U, self.I, self.K, self.ld, self.n,
                    learning_rate, weight_decay, clip_value, self.seed, self.epochs_done, self.inflight, stream))
        self.epochs_done += 1

    def download(self, W, H):
        import torch
        with torch.cuda.device(self.dev):
            _lib.download_factor(self.dW, self.K, W)
            _lib.download_factor(self.dH, self.K, H)

    def dense_f64(self):
        import torch
        out = []
        with torch.cuda.device(self.dev):
            for m in (self.dW, self.dH):
                t = torch.empty((m.shape[0], self.K), dtype=torch.float64, device=self.dev)
                _lib.check(self._L.cymf_unpack_rows_dev(_lib.ptr(m), _lib.ptr(t), self.dtype, m.shape[0], self.K,
                                                        self.ld, _lib.stream_ptr()))
                out.append(t)
        return tuple(out)

    def snapshot(self):
        return self.dW.clone(), self.dH.clone()

    def restore(self, snap):
        self.dW.copy_(snap[0])
        self.dH.copy_(snap[1])

    @property
    def bytes_per_update(self):
        """Algorithmic bytes per sample: read+write of 2 rows (x state copies) + the propensity."""
        es = 4 if self.dtype == _lib.F32 else 8
        copies = {_lib.SGD: 1, _lib.ADAGRAD: 2, _lib.ADAM: 3}[self.opt]
        return 4 * self.K * es * copies + es
