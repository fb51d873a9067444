"""`cymf.BPR` on a B200: same constructor, `fit` signature, attributes and error behaviour as the reference
class (cymf/bpr.pyx:37-190); the prange loop of `_fit_bpr` (bpr.pyx:160-171) runs as CUDA kernels
(cymf_b200/csrc/bpr.cu) reached through the C ABI of include/cymf_b200.h.

Host logic kept verbatim in Python/NumPy, as in the reference: input coercion (bpr.pyx:78-87), seeded
init (bpr.pyx:97-101), the one-time shuffle (bpr.pyx:104), per-epoch validation and early stopping
(bpr.pyx:173-190).  Everything inside the epoch loop is on the device; there is no CPU fallback.

Keyword-only extras (not in the reference):
    mode   "hogwild" (default): warp-group-per-triplet kernel, Philox negatives, `dtype` storage;
           "replay": serialized f64 kernel driven by the reference's own mt19937 negative stream --
           reproduces `num_threads=1` results of the reference step by step.
    dtype  "float32" (default, 16-byte vector gathers) or "float64" for the Hogwild kernel.
    scatter "auto" | "store" | "red": how SGD updates are written back (vector stores vs red.global.add).
"""
import ctypes as C

import numpy as np
from scipy import sparse
from sklearn import utils

from . import _lib
from .host import init_missing_factors, run_epochs


def _tqdm(total, verbose, ncols=120):
    from tqdm import tqdm
    return tqdm(total=total, leave=True, ncols=ncols, disable=not verbose)


class BPR(object):
    """
    Bayesian Personalized Ranking (BPR), https://arxiv.org/pdf/1205.2618.pdf

    Attributes:
        num_components (int): A dimensionality of latent vector
        learning_rate (double): A learning rate
        optimizer (str): 'adam', 'adagrad' or 'sgd'
        weight_decay (double): A coefficient of weight decay
        W (np.ndarray[double, ndim=2]): User latent vectors
        H (np.ndarray[double, ndim=2]): Item latent vectors
    """

    def __init__(self, num_components=20, learning_rate=0.001, optimizer="adam", weight_decay=0.01, *,
                 mode="hogwild", dtype="float32", scatter="auto", seed=1234, max_inflight=None, device=None):
        self.num_components = int(num_components)
        self.learning_rate = float(learning_rate)
        self.optimizer = optimizer
        self.weight_decay = float(weight_decay)
        self.W = None
        self.H = None
        if self.optimizer not in ("sgd", "adagrad", "adam"):
            raise Exception(f"{self.optimizer} is invalid.")            # bpr.pyx:65-66
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
        self.n_applied_ = 0          # accepted (non-colliding) triplets of the last fit
        self.n_attempted_ = 0

    def fit(self, X, num_epochs=10, num_threads=1, valid_evaluator=None, early_stopping=False, verbose=True):
        """
        Training BPR model with Gradient Descent.

        Args:
            X: A user-item interaction matrix.
            num_epochs (int): A number of epochs.
            num_threads (int): accepted for signature compatibility; the GPU grid replaces the thread pool.
            verbose (bool): Whether to show the progress of training.
        """
        if X is None:
            raise ValueError()
        if sparse.isspmatrix(X):
            X = X.tocsr()
        elif isinstance(X, np.ndarray):
            X = sparse.csr_matrix(X)
        else:
            raise ValueError()
        X = X.astype(np.float64, copy=False)     # the reference copies (wmf.pyx:84 / bpr.pyx:87); X is only read here

        self.valid_evaluator = valid_evaluator
        self.valid_dcg = -np.inf
        self.count = 0
        self.early_stopping = early_stopping
        if early_stopping and self.valid_evaluator is None:
            raise ValueError()

        init_missing_factors(self, X.shape[0], X.shape[1])               # bpr.pyx:97-101

        users, positives = utils.shuffle(*(X.nonzero()))
        return self._fit_bpr(users.astype(np.int32), positives.astype(np.int32), X, num_epochs,
                             self.learning_rate, self.weight_decay, num_threads, verbose)

    # ------------------------------------------------------------------------------------------------------
    def _fit_bpr(self, users, positives, X, num_epochs, learning_rate, weight_decay, num_threads, verbose):
        """Device replacement of `BPR._fit_bpr` (bpr.pyx:117-190); same arguments."""
        self.W = np.ascontiguousarray(self.W, dtype=np.float64)
        self.H = np.ascontiguousarray(self.H, dtype=np.float64)
        if not hasattr(self, "valid_dcg"):
            self.valid_dcg = -np.inf
        # model.W / model.H are updated in place, like the reference's typed views (bpr.pyx:127-128)
        sess = BprSession(self.W, self.H, users, positives, X, self.optimizer, mode=self.mode, dtype=self.dtype,
                          scatter=self.scatter, seed=self.seed, max_inflight=self.max_inflight, device=self.device)
        run_epochs(self, sess, num_epochs, lambda: sess.epoch(learning_rate, weight_decay), verbose, ncols=120)
        self.n_applied_ = sess.applied()
        self.n_attempted_ = sess.N * sess.epochs_done


class BprSession(object):
    """Device-resident state of one `_fit_bpr` call: factors (+ optimizer state, rebuilt per fit as in
    bpr.pyx:149-156), the shuffled (user, positive) pairs and the CSR used for the membership test.
    `epoch()` enqueues one pass of the hot loop (bpr.pyx:162-169) on the current CUDA stream."""

    def __init__(self, W, H, users, positives, X, optimizer, *, mode="hogwild", dtype="float32", scatter="auto",
                 seed=1234, max_inflight=None, device=None):
        torch = _lib.require_cuda()
        self._L = _lib.lib()
        self.dev = dev = torch.device(device if device is not None else "cuda")
        self.U, self.I = X.shape
        self.K = K = W.shape[1]
        self.N = N = int(users.shape[0])
        X = X.tocsr()
        if not X.has_sorted_indices:
            X = X.sorted_indices()
        self.replay = mode == "replay"
        self.dtype = _lib.F64 if self.replay else _lib.DTYPES[dtype]
        self.opt = opt = _lib.OPTIMIZERS[optimizer]
        tdt = torch.float64 if self.dtype == _lib.F64 else torch.float32
        self.ld = ld = _lib.ld_for(K)
        U, I = self.U, self.I
        with torch.cuda.device(dev):
            # Upload order = order of first use: factors and CSR first, then the pair list in ranges on a copy
            # stream, so that the first epoch starts on range 0 while the later ranges are still crossing PCIe
            # (the kernel needs every factor row from its first triplet on, but only the pairs it has reached).
            self.d_users = torch.empty(N, dtype=torch.int32, device=dev)      # allocated first, filled last
            self.d_pos = torch.empty(N, dtype=torch.int32, device=dev)
            self.d_indptr = torch.from_numpy(X.indptr.astype(np.int64)).to(dev, non_blocking=True)
            self.d_indices = torch.from_numpy(np.ascontiguousarray(X.indices, np.int32)).to(dev, non_blocking=True)
            self.dW = _lib.upload_factor(W, self.dtype, dev)
            self.dH = _lib.upload_factor(H, self.dtype, dev)
            src_u = torch.from_numpy(np.ascontiguousarray(users, np.int32))
            src_p = torch.from_numpy(np.ascontiguousarray(positives, np.int32))
            self._ranges = []
            n_ranges = 4 if (N >= (1 << 22) and mode != "replay") else 1
            copy_stream = torch.cuda.Stream(device=dev) if n_ranges > 1 else torch.cuda.current_stream()
            bounds = [N * q // n_ranges for q in range(n_ranges + 1)]
            if n_ranges > 1:
                # the buffers were just handed out by the caching allocator in main-stream order: earlier main-stream
                # work may still be using their memory, so the copy stream starts behind it (which is also the
                # upload order wanted: factors and CSR first)
                copy_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(copy_stream):
                for a, b in zip(bounds[:-1], bounds[1:]):
                    self.d_users[a:b].copy_(src_u[a:b], non_blocking=True)
                    self.d_pos[a:b].copy_(src_p[a:b], non_blocking=True)
                    if n_ranges > 1:
                        ev = torch.cuda.Event()
                        ev.record(copy_stream)
                        self._ranges.append((a, b, ev))
            self._keep_src = (src_u, src_p)                      # alive until the copies have been consumed
            self.state = []
            if opt == _lib.ADAGRAD:
                self.state = [torch.ones((U, ld), dtype=tdt, device=dev), torch.ones((I, ld), dtype=tdt, device=dev)]
            elif opt == _lib.ADAM:
                self.state = [torch.zeros((n, ld), dtype=tdt, device=dev) for n in (U, I, U, I)]
            self.d_applied = torch.zeros(1, dtype=torch.int64, device=dev)
        sp = [_lib.ptr(t) for t in self.state] + [None] * (4 - len(self.state))
        self.f = _lib.Factors(_lib.ptr(self.dW), _lib.ptr(self.dH), sp[0], sp[1], sp[2], sp[3])
        # "auto": 128-bit red.global.add.v4.f32 costs the same as a vector store (measured) and loses no update --
        # parameters and optimizer state alike; f64 reductions are scalar and 2.4x slower than stores, so f64
        # keeps plain stores.
        self.scatter = {"auto": 1 if self.dtype == _lib.F32 else 0, "store": 0, "red": 1}[scatter]
        # bounds Hogwild staleness on small matrices; large ones fill the machine
        self.inflight = int(max_inflight) if max_inflight is not None else max(1024, N // 256)
        self.seed = int(seed)
        self.gen = _lib.HostRng(1234) if self.replay else None   # bpr.pyx:141: one generator per fit
        self.epochs_done = 0
        # bytes the host hands to / takes from the device for this fit (bench.py's e2e accounting)
        self.h2d_bytes = (W.nbytes + H.nbytes + 2 * 4 * N + 8 * (U + 1) + 4 * X.indices.shape[0])
        self.d2h_bytes = W.nbytes + H.nbytes + 8

    def epoch(self, learning_rate, weight_decay):
        import torch
        L, f = self._L, self.f
        with torch.cuda.device(self.dev):
            stream = _lib.stream_ptr()
            if self.replay:
                neg = torch.from_numpy(self.gen.below(self.I, self.N)).to(self.dev)
                _lib.check(L.cymf_bpr_replay_epoch_dev(
                    C.byref(f), self.opt, _lib.ptr(self.d_users), _lib.ptr(self.d_pos), _lib.ptr(neg), self.N,
                    _lib.ptr(self.d_indptr), _lib.ptr(self.d_indices), self.U, self.I, self.K, self.ld,
                    learning_rate, weight_decay, _lib.ptr(self.d_applied), stream))
                self._keep = neg                                 # alive until the kernel has consumed it
            else:
                # the first epoch follows the pair ranges as they arrive; the negatives are keyed by the pair's
                # position in the epoch, so the result does not depend on how the epoch is cut
                ranges = self._ranges if (self.epochs_done == 0 and self._ranges) else [(0, self.N, None)]
                for a, b, ev in ranges:
                    if ev is not None:
                        torch.cuda.current_stream().wait_event(ev)
                    if b > a:
                        _lib.check(L.cymf_bpr_hogwild_range_dev(
                            C.byref(f), self.dtype, self.opt, self.scatter, _lib.ptr(self.d_users[a:]),
                            _lib.ptr(self.d_pos[a:]), b - a, _lib.ptr(self.d_indptr), _lib.ptr(self.d_indices),
                            self.U, self.I, self.K, self.ld, learning_rate, weight_decay, self.seed,
                            self.epochs_done, self.inflight, _lib.ptr(self.d_applied), a, stream))
                self._ranges = []
        self.epochs_done += 1

    def download(self, W, H):
        import torch
        with torch.cuda.device(self.dev):
            _lib.download_factor(self.dW, self.K, W)
            _lib.download_factor(self.dH, self.K, H)

    def dense_f64(self):
        """(W, H) as dense float64 DEVICE tensors [rows, K] -- what the on-device evaluator scores."""
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
        """Device copy of the factors (best-epoch bookkeeping of early stopping, bpr.pyx:182-183)."""
        return self.dW.clone(), self.dH.clone()

    def restore(self, snap):
        self.dW.copy_(snap[0])
        self.dH.copy_(snap[1])

    def applied(self):
        return int(self.d_applied.item())

    @property
    def bytes_per_update(self):
        """Algorithmic bytes per APPLIED update: read+write of 3 rows (x state copies) + (u, i)."""
        es = 4 if self.dtype == _lib.F32 else 8
        copies = {_lib.SGD: 1, _lib.ADAGRAD: 2, _lib.ADAM: 3}[self.opt]
        return 6 * self.K * es * copies + 8
