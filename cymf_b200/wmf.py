"""`cymf.WMF` on B200s: same constructor / `fit` / `_als` signatures and attributes as the reference class
(cymf/wmf.pyx:32-174); the Gram product and the prange row loop of `_als` (wmf.pyx:142-174) run as CUDA kernels
(cymf_b200/csrc/: als_ws.cu -- warp-specialised tensor-core row solver --, als_dual.cu -- short rows in their dual
form --, als_tc.cu, als.cu, tc_gemm*.cu) reached through the C ABI of include/cymf_b200.h.

Host logic kept in Python as in the reference: input coercion and seeded init (wmf.pyx:69-92), the epoch loop
with per-epoch validation / early stopping (wmf.pyx:110-132).  The reference re-transposes X twice per epoch
(wmf.pyx:112); here both orientations are built once per fit.

The reference factorises a dense K x K matrix per row with LAPACK dgesv; here each row is solved by conjugate
gradient to a relative residual `cg_tol` (default 1e-6 in float32, 1e-10 in float64), which keeps W and H
within 1e-4 (relative) of the reference's -- see tests/test_wmf_gpu.py.  In float32 with K <= 128 the row's K x K
matrix (wmf.pyx:161-166) is built once on the tensor cores from ONE gather of the row (rows of at most 64 entries: the
n x n matrix of the dual system instead) and the iteration runs out of registers; float64 and K > 128 stream the row
once per iteration (cymf_als_cg_dev).

Environment switches (A/B runs, tests): CYMF_ALS_WS=0 (CTA-per-row solver instead of the warp-specialised one),
CYMF_ALS_DUAL=0 / CYMF_ALS_DUAL_MAX, CYMF_ALS_WS_TMA=0|1 (cp.async / TMA gather), CYMF_ALS_SHORT, CYMF_ALS_ROWS=cg,
CYMF_NO_TCGEN05=1, CYMF_ALS_TAIL_DIVISOR, CYMF_ALS_GRAPH=0, CYMF_CHOL_BLOCKED=0.

Multi-GPU (one process per GPU, torch.distributed/NCCL already initialised): rows of each half sweep are
partitioned over the ranks (heaviest-first round-robin deal, so row counts are equal and nnz is balanced), every
rank holds the whole fixed side in the coordinates of its Cholesky change of variables, solves its own block, then
writes the transformed block into every rank's copy from the GEMM epilogue (NVLink peer stores) and the K x K Gram
partials of the freshly solved blocks are summed by peer loads.  world_size == 1 runs the same code.
"""
import ctypes as C
import os

import numpy as np
from scipy import sparse

from . import _lib
from .host import init_missing_factors, run_epochs


class WMF(object):
    """
    Weighted Matrix Factorization (WMF), http://yifanhu.net/PUB/cf.pdf

    Attributes:
        num_components (int): A dimensionality of latent vector
        weight_decay (double): A coefficient of weight decay
        weight (double): A weight for positive feedbacks.
        W (np.ndarray[double, ndim=2]): User latent vectors
        H (np.ndarray[double, ndim=2]): Item latent vectors
    """

    def __init__(self, num_components=20, weight_decay=0.01, weight=10.0, *, dtype="float32", cg_tol=None,
                 cg_max_iter=None, device=None, distributed=False, peer_gather=True, solver="transformed",
                 prep="device", heavy_min=4096):
        self.num_components = int(num_components)
        self.weight_decay = float(weight_decay)
        self.weight = float(weight)
        self.W = None
        self.H = None
        if dtype not in _lib.DTYPES:
            raise ValueError("dtype must be 'float32' or 'float64'")
        if self.num_components > 256:
            raise ValueError("cymf_b200.WMF supports num_components <= 256")
        self.dtype = dtype
        self.cg_tol = cg_tol
        self.cg_max_iter = cg_max_iter
        self.device = device
        self.distributed = distributed
        self.peer_gather = peer_gather
        self.solver = solver
        if prep not in ("device", "host"):
            raise ValueError("prep must be 'device' or 'host'")
        self.prep = prep                 # where X^T, the row deal and the relabelled blocks are built
        self.heavy_min = int(heavy_min)  # rows with at least this many entries are solved directly (0 = never)
        self.cg_iterations_ = 0          # CG iterations summed over rows, last fit
        self.cg_unconverged_ = 0         # rows that stopped at cg_max_iter, last fit

    def fit(self, X, num_epochs=5, num_threads=1, valid_evaluator=None, early_stopping=False, verbose=True):
        """
        Training WMF model with ALS.

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
        init_missing_factors(self, X.shape[0], X.shape[1])               # wmf.pyx:88-92
        self._fit_als(X, num_epochs, num_threads, verbose)

    def _tolerances(self):
        tol = self.cg_tol if self.cg_tol is not None else (1e-6 if self.dtype == "float32" else 1e-10)
        iters = self.cg_max_iter if self.cg_max_iter is not None else 2 * self.num_components
        return float(tol), int(iters)

    def _fit_als(self, X, num_epochs, num_threads, verbose):
        """Device replacement of `WMF._fit_als` (wmf.pyx:97-132)."""
        self.W = np.ascontiguousarray(self.W, dtype=np.float64)
        self.H = np.ascontiguousarray(self.H, dtype=np.float64)
        if not hasattr(self, "valid_dcg"):
            self.valid_dcg = -np.inf
        tol, iters = self._tolerances()
        sess = AlsSession(X, self.W, self.H, self.weight_decay, self.weight, dtype=self.dtype, cg_tol=tol,
                          cg_max_iter=iters, device=self.device, distributed=self.distributed,
                          peer_gather=self.peer_gather, solver=self.solver, prep=self.prep,
                          heavy_min=self.heavy_min)
        # how solved blocks reach the other ranks: "single" | "peer-store" (fused into the GEMM epilogue) | "nccl"
        self.gather_mode_ = "peer-store" if sess.peer else ("nccl" if sess.dist else "single")
        self.row_solver_ = sess.row_solver
        run_epochs(self, sess, num_epochs, sess.epoch, verbose, ncols=100)
        self.cg_iterations_, self.cg_unconverged_ = sess.stats()
        self.transfer_bytes_ = (sess.h2d_bytes, sess.d2h_bytes)     # host<->device bytes this fit moved

    def _als(self, indptr, indices, X, Y, num_threads=1):
        """`WMF._als(indptr, indices, X, Y, num_threads)` (wmf.pyx:136): solves every row of the HOST array X in
        place from the HOST array Y; goes through the host-buffer C entry point."""
        _lib.require_cuda()
        if not (isinstance(X, np.ndarray) and X.dtype == np.float64 and X.flags.c_contiguous):
            raise ValueError("X must be a C-contiguous float64 array")
        Y = np.ascontiguousarray(Y, dtype=np.float64)
        ip = np.ascontiguousarray(indptr, np.int32)
        ix = np.ascontiguousarray(indices, np.int32)
        tol, iters = self._tolerances()
        done = C.c_int64(0)
        p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        _lib.check(_lib.lib().cymf_als_half_host(p(ip), p(ix), p(X), p(Y), X.shape[0], Y.shape[0], X.shape[1],
                                                 self.weight_decay, self.weight, _lib.DTYPES[self.dtype], tol, iters,
                                                 C.byref(done)))
        self.cg_iterations_ = int(done.value)


def _deal(degrees, world):
    """Heaviest-first round-robin deal of rows to ranks.  Returns (slot -> old row id or -1, rows per rank):
    rank r owns the contiguous slots [r*R, (r+1)*R), sorted by decreasing degree, phantoms (-1) at the end."""
    n = degrees.shape[0]
    by_weight = np.argsort(-degrees, kind="stable")
    R = (n + world - 1) // world
    slots = np.full(R * world, -1, dtype=np.int64)
    for r in range(world):
        mine = by_weight[r::world]
        slots[r * R:r * R + mine.shape[0]] = mine
    return slots, R


def _relabel(Xcsr, row_slots, col_new_index, n_cols):
    """CSR whose row q is old row row_slots[q] (empty for phantoms) with columns renamed by col_new_index."""
    ext = sparse.vstack([Xcsr, sparse.csr_matrix((1, Xcsr.shape[1]))]).tocsr()
    rows = np.where(row_slots >= 0, row_slots, Xcsr.shape[0])
    P = ext[rows]
    return sparse.csr_matrix((P.data, col_new_index[P.indices].astype(np.int32), P.indptr),
                             shape=(rows.shape[0], n_cols))


# Symmetric (NVLink peer-mapped) buffers are expensive to set up (an allocation + a rendezvous of all ranks, ~0.1 s
# each), and `WMF.fit` builds one session per call: buffers are kept per (name, shape, dtype, device, world) and reused
# by later sessions of the same process group.  Every element is rewritten before it is read (publish step), so no
# stale content can leak between fits.
_SYMM_CACHE = {}


class AlsSession(object):
    """Device-resident state of one `_fit_als` call (both CSR orientations of this rank's row blocks, full
    replicas of W and H in dealt order); `epoch()` = user half sweep + item half sweep (wmf.pyx:111-112)."""

    def __init__(self, X, W, H, weight_decay, weight, *, K=None, dtype="float32", cg_tol=1e-6, cg_max_iter=128, device=None,
                 distributed=False, stage_rows=0, force_width=0, solver="transformed", peer_gather=True,
                 prep="device", overlap_classes=True, heavy_min=4096):
        torch = _lib.require_cuda()
        self.heavy_min = int(heavy_min)
        self.tail_divisor = int(os.environ.get("CYMF_ALS_TAIL_DIVISOR", "256"))     # tuning hook (tools/c5_als.py)
        # rows of at most this many entries go to the streaming CG kernel even when the tensor-core solver is on: its
        # cost grows with the row length (~145 SM-clocks per entry) while the one-pass solver pays ~10 k SM-clocks per
        # row whatever its length (K x K CG from registers), so very short rows are cheaper streamed.  0 = never.
        # Measured, ml-20m shape K=128 (gpurun_out r2_sweep_v5): 13.1 ms/epoch at 0, 12.15 at 96 and 128, 12.4 at 192.
        self.short_max = int(os.environ.get("CYMF_ALS_SHORT", "112"))
        # Rows of at most dual_max entries are solved in their dual (n x n) form, several rows per 128-slot tensor-core
        # tile (cymf_als_rows_dual_dev); set per session below: min(128, ld) rounded down to a tile class.  When it is
        # on, the streaming kernel is not used at all: longer rows take the one-pass K x K solver.  CYMF_ALS_DUAL=0: off.
        self.dual_max = 0
        # Long rows: the warp-specialised persistent solver (cymf_als_rows_ws_dev) with a host-side LPT schedule of the
        # rows over one CTA per SM; CYMF_ALS_WS=0 falls back to the CTA-per-row kernel (cymf_als_rows_tc_dev).
        self.use_ws = os.environ.get("CYMF_ALS_WS", "1") == "1"
        self.ws_row_cost = int(os.environ.get("CYMF_ALS_WS_ROWCOST", "256"))
        self._ws = {}
        self.peer_error = None
        self._unperm = {}
        self.force_width = int(force_width)
        if solver not in ("transformed", "pcg", "cg"):
            raise ValueError("solver must be 'transformed', 'pcg' or 'cg'")
        if int(K if K is not None else W.shape[1]) > 128:
            solver = "cg"      # the Cholesky change of variables and the tensor-core kernels stop at K = 128: plain CG
        self.solver = solver
        import torch.distributed as dist
        self._L = _lib.lib()
        # distributed=True: `fit` is a COLLECTIVE over the default process group -- every rank must call it with the
        # same X (the reference has no multi-device mode).  The default (False) keeps a process that
        # uses torch.distributed for something else -- e.g. independent per-rank fits -- free of hidden collectives.
        self.dist = dist if (distributed in ("auto", True) and dist.is_available() and dist.is_initialized()
                             and dist.get_world_size() > 1) else None
        self.world = self.dist.get_world_size() if self.dist else 1
        self.rank = self.dist.get_rank() if self.dist else 0
        self.dev = dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.dtype = _lib.DTYPES[dtype]
        self.tdt = tdt = torch.float32 if self.dtype == _lib.F32 else torch.float64
        self.K = K = int(K if K is not None else W.shape[1])
        # Row solver of the half sweep.  "tc": one pass over the row -- S = sum y~ y~^T on the tensor cores, CG out of
        # registers (cymf_als_rows_tc_dev; f32, K <= 128, transformed coordinates; rows are padded to a multiple of
        # 32 columns).  "cg": the streaming CG kernel (f64, or CYMF_ALS_ROWS=cg / CYMF_NO_TCGEN05=1 for A/B runs).
        self.row_solver = "cg"
        if (self.dtype == _lib.F32 and K <= 128 and solver == "transformed" and os.environ.get("CYMF_NO_TCGEN05") != "1"
                and os.environ.get("CYMF_ALS_ROWS", "tc") == "tc"):
            self.row_solver = "tc"
        self.ld = ld = (K + 31) // 32 * 32 if self.row_solver == "tc" else _lib.ld_for(K)
        if self.row_solver == "tc" and os.environ.get("CYMF_ALS_DUAL", "1") == "1":
            # measured, ml-20m shape K = 128: 11.0 ms / epoch with rows of 65..128 entries in the dual kernel, 9.3 ms
            # with those rows in the warp-specialised one-pass solver (its 128-slot tile class costs more than the K x K solve)
            cap = 128 if ld >= 128 else (64 if ld >= 64 else 32)         # largest tile class this ld supports
            self.dual_max = min(cap, int(os.environ.get("CYMF_ALS_DUAL_MAX", "64" if self.use_ws else "128")))
        self.wd, self.weight = float(weight_decay), float(weight)
        self.cg_tol, self.cg_max_iter, self.stage_rows = float(cg_tol), int(cg_max_iter), int(stage_rows)
        self.prep = prep
        self.overlap_classes = bool(overlap_classes)
        with torch.cuda.device(dev):
            if prep == "device":
                blk_u, blk_i = self._prepare_on_device(X)
            else:
                blk_u, blk_i = self._prepare_on_host(X)
            self.csr_u, self.csr_i = blk_u, blk_i
            self.block_nnz = (int(blk_u[1].numel()), int(blk_i[1].numel()))
            self.classes_u, self.heavy_u = self._classes(blk_u[0])
            self.classes_i, self.heavy_i = self._classes(blk_i[0])
            n_slabs = max(int(h[1][-1]) if h else 0 for h in (self.heavy_u, self.heavy_i))
            self.ws_heavy = torch.empty(max(1, int(self._L.cymf_als_heavy_workspace_doubles(n_slabs, K, ld))),
                                        dtype=torch.float64, device=dev)
            U, I, Up, Ip = self.U, self.I, self.slot_u.shape[0], self.slot_i.shape[0]
            self.order_u = torch.arange(self.Ru, dtype=torch.int32, device=dev)
            self.order_i = torch.arange(self.Ri, dtype=torch.int32, device=dev)
            self.dW = self._upload(W, self.slot_u)
            self.dH = self._upload(H, self.slot_i)
            nws = int(self._L.cymf_gram_workspace_doubles(max(Up, Ip), K))
            self.ws = torch.empty(max(nws, 1), dtype=torch.float64, device=dev)
            self.g64 = torch.empty(K * K, dtype=torch.float64, device=dev)
            self.G = torch.empty(ld * ld, dtype=tdt, device=dev)       # [ld, ld], zero padded
            self.Ginv = torch.empty(ld * ld, dtype=tdt, device=dev)
            self.By, self.Bfwd, self.Bbwd = (torch.empty(ld * ld, dtype=tdt, device=dev) for _ in range(3))
            self.queue = torch.zeros(8, dtype=torch.int32, device=dev)     # one work-queue head per row class
            self.d_stats = torch.zeros(2, dtype=torch.int64, device=dev)
            self.d_debug = torch.zeros(32, dtype=torch.int64, device=dev)  # first timed-out hand-over of the WS solver
            self.d_info = torch.zeros(1, dtype=torch.int32, device=dev)    # Cholesky pivot failures (checked in stats())
            # Transformed solver: every rank keeps the WHOLE fixed side in the coordinates y~ = L^-1 y (Yt["item"] is the
            # item factors as the user half sweep reads them, Yt["user"] the user factors for the item half sweep) and
            # its own block of each factor matrix in the original coordinates.  After a half sweep the solved block is
            # transformed for the NEXT half sweep and written into all ranks' Yt straight from the GEMM epilogue
            # (NVLink peer stores into symmetric memory): the all-gather of SURVEY.md 8(e) moves y~ instead of y, so no
            # rank ever transforms rows it does not own.  The K x K Gram partials meet in a symmetric buffer too.
            self.Yt = {"user": torch.zeros((Up, ld), dtype=tdt, device=dev),
                       "item": torch.zeros((Ip, ld), dtype=tdt, device=dev)}
            self.gpart = torch.zeros(K * K, dtype=torch.float64, device=dev)
            self.peer = None
            if self.dist and peer_gather and self.solver == "transformed":
                self.peer = self._make_symmetric()
            self._stale = {"user": False, "item": False}      # other ranks' blocks of dW / dH are out of date
            self._pub_side = None                             # side whose Yt / transforms were published last
            self._side_streams = [torch.cuda.Stream(device=dev) for _ in range(3)]
            self._graph, self.graph_error, self.graph_launches = None, None, 0
            self.use_graph = os.environ.get("CYMF_ALS_GRAPH", "1") == "1"
        self.epochs_done = 0
        self.kernel_events = None        # set to [] to record (algorithmic bytes, start, end) CUDA events per row-solver launch
        self.trace = None                # set to [] to record (label, event) at the phase boundaries (eager launches)
        self.h2d_bytes = self._nbytes(W) + self._nbytes(H) + 8 * (self.Ru + self.Ri + 2) + 4 * sum(self.block_nnz)
        self.d2h_bytes = self._nbytes(W) + self._nbytes(H)

    @staticmethod
    def _nbytes(a):
        return int(a.nbytes) if isinstance(a, np.ndarray) else int(a.numel() * a.element_size())

    def _prepare_on_host(self, X):
        """scipy / NumPy construction of this rank's row blocks (the reference's own tools: `X.T.tocsr()`,
        wmf.pyx:112).  Kept as the checker of the device path and for the gloo CPU tests."""
        import torch
        dev = self.dev
        X = X.tocsr()
        self.U, self.I = U, I = X.shape
        XT = X.T.tocsr()
        # dealt (permuted + padded) index spaces of users and items
        self.slot_u, self.Ru = _deal(np.diff(X.indptr), self.world)
        self.slot_i, self.Ri = _deal(np.diff(XT.indptr), self.world)
        new_u = np.empty(U, np.int64); new_u[self.slot_u[self.slot_u >= 0]] = np.flatnonzero(self.slot_u >= 0)
        new_i = np.empty(I, np.int64); new_i[self.slot_i[self.slot_i >= 0]] = np.flatnonzero(self.slot_i >= 0)
        Up, Ip = self.slot_u.shape[0], self.slot_i.shape[0]
        lo_u, lo_i = self.rank * self.Ru, self.rank * self.Ri
        blk_u = _relabel(X, self.slot_u[lo_u:lo_u + self.Ru], new_i, Ip)
        blk_i = _relabel(XT, self.slot_i[lo_i:lo_i + self.Ri], new_u, Up)
        self.nnz = int(X.nnz)
        self.d_slot_u, self.d_slot_i = (torch.from_numpy(a).to(dev) for a in (self.slot_u, self.slot_i))
        self.d_row_slot_u, self.d_row_slot_i = torch.from_numpy(new_u).to(dev), torch.from_numpy(new_i).to(dev)

        def up(a, dt):
            return torch.from_numpy(np.ascontiguousarray(a, dt)).to(dev, non_blocking=True)
        return (up(blk_u.indptr, np.int64), up(blk_u.indices, np.int32)), (up(blk_i.indptr, np.int64), up(blk_i.indices, np.int32))

    def _prepare_on_device(self, X):
        """The same construction with cymf_b200/csrc/prep.cu: the CSR is uploaded once (or already lives on the
        device: X = (indptr int64 tensor, indices int32 tensor, (U, I))), transpose / deal / relabel run as kernels."""
        import torch
        from . import prep
        dev = self.dev
        if isinstance(X, tuple):
            ip, ix, (U, I) = X
            ip, ix = ip.to(dev), ix.to(dev)
        else:
            X = X.tocsr()
            U, I = X.shape
            ip = torch.from_numpy(X.indptr.astype(np.int64)).to(dev)
            ix = torch.from_numpy(np.ascontiguousarray(X.indices, np.int32)).to(dev)
        self.U, self.I = int(U), int(I)
        self.nnz = int(ix.numel())
        t_ip, t_ix = prep.transpose_csr(ip, ix, self.U, self.I)
        slot_u, row_slot_u, self.Ru = prep.deal_rows(ip, self.U, self.world)
        slot_i, row_slot_i, self.Ri = prep.deal_rows(t_ip, self.I, self.world)
        lo_u, lo_i = self.rank * self.Ru, self.rank * self.Ri
        blk_u = prep.csr_block(ip, ix, slot_u[lo_u:lo_u + self.Ru], row_slot_i)
        blk_i = prep.csr_block(t_ip, t_ix, slot_i[lo_i:lo_i + self.Ri], row_slot_u)
        self.d_slot_u, self.d_slot_i = slot_u, slot_i
        self.d_row_slot_u, self.d_row_slot_i = row_slot_u, row_slot_i
        self.slot_u, self.slot_i = slot_u.cpu().numpy(), slot_i.cpu().numpy()
        return blk_u, blk_i

    def _make_symmetric(self):
        """Move Yt["user"], Yt["item"] and the Gram partial into NVLink-mapped symmetric memory.  Returns
        {name: (handle, [base pointer of every rank's copy])} or None (then NCCL collectives carry the exchange)."""
        import torch
        moved = []
        try:
            import torch.distributed._symmetric_memory as symm
            group = self.dist.group.WORLD
            gname = group.group_name if hasattr(group, "group_name") else group
            for name, old in (("user", self.Yt["user"]), ("item", self.Yt["item"]), ("gram", self.gpart)):
                key = (name, tuple(old.shape), old.dtype, str(self.dev), self.world, str(gname))
                if key not in _SYMM_CACHE:
                    new = symm.empty(tuple(old.shape), dtype=old.dtype, device=self.dev)
                    hdl = symm.rendezvous(new, gname)
                    ptrs = [int(p) for p in hdl.buffer_ptrs]
                    if len(ptrs) != self.world or ptrs[self.rank] != new.data_ptr():
                        raise RuntimeError("unexpected symmetric-memory layout")
                    _SYMM_CACHE[key] = (new, hdl, ptrs)
                new, hdl, ptrs = _SYMM_CACHE[key]
                new.copy_(old)
                moved.append((name, new, hdl, ptrs))
            ok = torch.ones(1, device=self.dev)
        except Exception as exc:                                   # noqa: BLE001 - any failure -> NCCL path
            import warnings
            self.peer_error = repr(exc)
            warnings.warn(f"cymf_b200.WMF: symmetric (NVLink peer) memory unavailable, blocks are exchanged with NCCL "
                          f"collectives instead: {self.peer_error}", RuntimeWarning)
            ok = torch.zeros(1, device=self.dev)
        self.dist.all_reduce(ok, op=self.dist.ReduceOp.MIN)        # all ranks take the same path
        if ok.item() < 1:
            return None
        out = {}
        for name, new, hdl, ptrs in moved:
            if name == "gram":
                self.gpart = new
            else:
                self.Yt[name] = new
            out[name] = (hdl, ptrs)
        return out

    def _upload(self, host, slots):
        """Dense f64 [rows, K] ndarray (or a device tensor [rows, ld] of the session dtype) -> device
        [len(slots), ld] in dealt order (phantom rows zero)."""
        import torch
        if isinstance(host, np.ndarray) and self.prep != "device":
            dealt = np.zeros((slots.shape[0], host.shape[1]), np.float64)
            dealt[slots >= 0] = host[slots[slots >= 0]]
            return _lib.upload_factor(dealt, self.dtype, self.dev, self.ld)
        t = _lib.upload_factor(host, self.dtype, self.dev, self.ld) if isinstance(host, np.ndarray) else host
        if t.shape[1] < self.ld:                                   # device tensor narrower than the session's row stride
            t = torch.cat([t, torch.zeros((t.shape[0], self.ld - t.shape[1]), dtype=t.dtype, device=t.device)], 1)
        d_slots = self.d_slot_u if slots is self.slot_u else self.d_slot_i
        t = torch.cat([t, torch.zeros((1, t.shape[1]), dtype=t.dtype, device=t.device)])     # row for the phantoms
        return t.index_select(0, torch.where(d_slots >= 0, d_slots, t.shape[0] - 1))

    # one half sweep: solve `rows_side` from the fixed side (wmf.pyx:136-174)
    def _classes(self, blk_indptr):
        """Split of the block's rows (already heaviest first): ((rows solved by CG with 16 warps, with 8, with 4),
        heavy) where heavy = (n_heavy, first_slab device int32[n_heavy+1]) are the leading rows of at least
        `heavy_min` entries that are solved directly (cymf_als_heavy_rows_dev), or None."""
        import torch
        lengths = np.ascontiguousarray(np.diff(blk_indptr.cpu().numpy()), np.int64)
        n = int(lengths.shape[0])
        if self.force_width:                                   # tuning hook: one CTA width for every row
            return {16: (n, 0, 0), 8: (0, n, 0), 4: (0, 0, n)}[self.force_width], None
        n16, n8 = C.c_int64(0), C.c_int64(0)
        _lib.check(self._L.cymf_als_row_classes(lengths.ctypes.data_as(C.c_void_p), n, self.dtype,
                                                self.ld, C.byref(n16), C.byref(n8)))
        b16, b8 = int(n16.value), int(n16.value + n8.value)
        heavy, nh = None, 0
        if (self.heavy_min > 0 and self.solver == "transformed" and self.dtype == _lib.F32 and self.ld % 32 == 0
                and self.ld <= 128 and os.environ.get("CYMF_NO_TCGEN05") != "1"):
            # Machine time is about the same either way (CG: ~0.05-0.11 us x n on ONE SM; direct: n / 512 slabs of
            # ~64 us spread over all SMs + a 0.35 ms solve), so the direct path only pays for rows that would stick
            # out as a tail: row time >= ~half of the block's time (~0.45 ns per entry)  <=>  n >= block nnz / 256.
            # Measured: ml-20m on one GPU has no such row (19.4 ms/epoch either way, 22.2 ms with a floor of 4096);
            # sharded over 8 GPUs the same rows are 8x larger relative to their block and do qualify.
            floor = max(self.heavy_min, int(lengths.sum()) // self.tail_divisor)
            if self.row_solver == "tc" and self.use_ws and "CYMF_ALS_TAIL_DIVISOR" not in os.environ:
                # With the warp-specialised solver a long row costs its CTA ~25 ns per entry (1600 cycles per 32-entry
                # chunk) and only delays the half sweep by what exceeds the CTA's fair share, while the direct path
                # costs ~0.9 ms per half sweep as soon as ONE row takes it (slab Gram + f64 LDL^T + stream join).
                # Measured, ml-20m shape on 4 GPUs (tools/als_tail_ab.py): 4.52 ms/epoch with 6 direct rows per rank,
                # 4.36 with 1, 3.44 with none (the 31 k-entry row then runs 0.8 ms in one CTA).  So: direct solve only
                # for a row that exceeds its CTA's fair share by more than ~1 ms = 40 k entries.
                floor = max(self.heavy_min, int(lengths.sum()) // int(self._L.cymf_als_ws_ctas()) + 40_000)
            while True:                                        # keep the slab workspace under 8 GB
                nh = int((lengths >= floor).sum())             # a prefix: lengths are sorted in decreasing order
                slabs = int(((lengths[:nh] + 511) // 512).sum())
                if slabs * (self.K * self.K + self.ld) * 8 <= 8e9:
                    break
                floor *= 2
            if nh:
                first = np.concatenate([[0], np.cumsum((lengths[:nh] + 511) // 512)]).astype(np.int32)
                heavy = (nh, first, torch.from_numpy(first).to(self.dev))
        self._n_long = getattr(self, "_n_long", {})
        self._n_long[id(blk_indptr)] = int((lengths > self.short_max).sum()) if self.short_max > 0 else n
        self._n_dual = getattr(self, "_n_dual", {})
        if self.dual_max > 0:                                  # lengths decrease: the dual classes are the tail of the block
            over = int((lengths > self.dual_max).sum())
            over64 = max(over, int((lengths > 64).sum()))
            over32 = max(over, int((lengths > 32).sum()))
            self._n_long[id(blk_indptr)] = over
            self._n_dual[id(blk_indptr)] = (over64 - over, over32 - over64, n - over32)
        if self.row_solver == "tc" and self.use_ws:
            first, last = nh, max(self._n_long[id(blk_indptr)], nh)      # block rows [first, last) take the one-pass solver
            if last > first:
                n_ctas = int(self._L.cymf_als_ws_ctas())
                ip = np.ascontiguousarray(blk_indptr.cpu().numpy(), np.int64)
                rows = np.arange(first, last, dtype=np.int32)
                cta_ptr = np.empty(n_ctas + 1, np.int32)
                rowinfo = np.empty(4 * (last - first), np.int32)
                _lib.check(self._L.cymf_als_ws_schedule_host(ip.ctypes.data_as(C.c_void_p), rows.ctypes.data_as(C.c_void_p),
                                                             last - first, n_ctas, self.ws_row_cost,
                                                             cta_ptr.ctypes.data_as(C.c_void_p),
                                                             rowinfo.ctypes.data_as(C.c_void_p)))
                self._ws[id(blk_indptr)] = (torch.from_numpy(rowinfo).to(self.dev), torch.from_numpy(cta_ptr).to(self.dev),
                                            n_ctas)
        return (max(b16 - nh, 0), max(b8 - max(nh, b16), 0), n - max(nh, b8)), heavy

    def _side(self, side):
        if side == "user":
            return self.dW, self.Ru, self.csr_u, self.order_u, self.classes_u, self.heavy_u, "item"
        return self.dH, self.Ri, self.csr_i, self.order_i, self.classes_i, self.heavy_i, "user"

    def _mark(self, label):
        """Dev hook (tools/als_phases.py): CUDA event at a phase boundary of the eagerly launched half sweep."""
        if self.trace is not None:
            import torch
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.trace.append((label, ev))

    def _publish(self, side):
        """This rank's block of `side`'s factor matrix has just been solved (original coordinates).  Make it the
        fixed side of the next half sweep: G = X^T X + wd I (wmf.pyx:142-143) from the ranks' K x K partials,
        G = L L^T, and X~ = X L^-T for this rank's rows written into EVERY rank's Yt[side]."""
        L, K, ld = self._L, self.K, self.ld
        X_full, R = self._side(side)[:2]
        x_blk = X_full[self.rank * R:(self.rank + 1) * R]
        stream = _lib.stream_ptr()
        es = 4 if self.dtype == _lib.F32 else 8
        self._mark(side + ":publish start")
        if not self.dist:
            _lib.check(L.cymf_gram_dev(_lib.ptr(x_blk), self.dtype, R, K, ld, self.wd, 1, _lib.ptr(self.ws),
                                       self.ws.numel(), _lib.ptr(self.g64), None, stream))
        else:
            _lib.check(L.cymf_gram_dev(_lib.ptr(x_blk), self.dtype, R, K, ld, self.wd, 0, _lib.ptr(self.ws),
                                       self.ws.numel(), _lib.ptr(self.gpart), None, stream))
            if self.peer is not None:
                # Gram all-reduce as peer loads: barrier (all partials written), every rank sums them in rank order
                hdl, ptrs = self.peer["gram"]
                self._mark(side + ":gram partial")
                hdl.barrier(channel=0)
                self._mark(side + ":barrier 1")
                parts = (C.c_void_p * self.world)(*ptrs)
                _lib.check(L.cymf_gram_sum_dev(parts, self.world, K, self.wd, _lib.ptr(self.g64), stream))
            else:
                self.dist.all_reduce(self.gpart)
                parts = (C.c_void_p * 1)(self.gpart.data_ptr())
                _lib.check(L.cymf_gram_sum_dev(parts, 1, K, self.wd, _lib.ptr(self.g64), stream))
        self._mark(side + ":gram sum")
        _lib.check(L.cymf_chol_transforms_dev(_lib.ptr(self.g64), K, ld, 0.0, self.dtype, _lib.ptr(self.By),
                                              _lib.ptr(self.Bfwd), _lib.ptr(self.Bbwd), _lib.ptr(self.d_info), stream))
        self._mark(side + ":cholesky")
        yt = self.Yt[side]
        off = self.rank * R * ld * es
        if self.peer is not None:
            hdl, ptrs = self.peer[side]
            outs = (C.c_void_p * self.world)(*[p + off for p in ptrs])
            _lib.check(L.cymf_rows_times_matrix_multi_dev(_lib.ptr(x_blk), outs, self.world, _lib.ptr(self.By),
                                                          self.dtype, R, ld, stream))
            self._mark(side + ":transform + peer stores")
            hdl.barrier(channel=0)                               # every rank's rows have landed in this rank's Yt
            self._mark(side + ":barrier 2")
        else:
            y_blk = yt[self.rank * R:(self.rank + 1) * R]
            _lib.check(L.cymf_rows_times_matrix_dev(_lib.ptr(x_blk), _lib.ptr(y_blk), _lib.ptr(self.By), self.dtype,
                                                    R, ld, stream))
            if self.dist:
                self.dist.all_gather_into_tensor(yt, y_blk)

    def _solve(self, side):
        """Solve this rank's block of `side` against the published fixed side (wmf.pyx:150-168)."""
        import torch
        L, K, ld = self._L, self.K, self.ld
        X_full, R, csr, order, classes, heavy, other = self._side(side)
        x_blk = X_full[self.rank * R:(self.rank + 1) * R]
        yt = self.Yt[other]
        es = 4 if self.dtype == _lib.F32 else 8
        stream = _lib.stream_ptr()
        # warm start in the coordinates x~ = L^T x of the fixed side's Cholesky factor
        self._mark(side + ":solve start")
        _lib.check(L.cymf_rows_times_matrix_dev(_lib.ptr(x_blk), _lib.ptr(x_blk), _lib.ptr(self.Bfwd), self.dtype, R, ld, stream))
        self._mark(side + ":warm start")
        main = torch.cuda.current_stream()
        start, joins, fork = 0, [], None
        if heavy:
            # the few very long rows: K x K matrix built once by MANY CTAs on the tensor cores, solved directly
            nh, first_host, first_dev = heavy
            st = self._side_streams[2] if self.overlap_classes else main
            if st is not main:
                fork = torch.cuda.Event()
                fork.record(main)
                st.wait_event(fork)
            with torch.cuda.stream(st):
                _lib.check(L.cymf_als_heavy_rows_dev(_lib.ptr(csr[0]), _lib.ptr(csr[1]), _lib.ptr(order), nh,
                                                     _lib.ptr(first_dev), int(first_host[-1]), _lib.ptr(x_blk),
                                                     _lib.ptr(yt), None, 0.0, self.dtype, K, ld, self.weight,
                                                     _lib.ptr(self.ws_heavy), self.ws_heavy.numel(), _lib.stream_ptr()))
                if st is not main:
                    done = torch.cuda.Event()
                    done.record(st)
                    joins.append(done)
            start = nh
        if self.row_solver == "tc":
            n_rows = start + sum(classes)
            n_long = max(self._n_long[id(csr[0])], start)     # rows [start, n_long) one-pass, [n_long, n_rows) streamed
            count = n_long - start
            if n_rows > n_long and self.dual_max > 0:
                pass                                          # launched after the long rows, below
            elif n_rows > n_long:
                st = self._side_streams[0] if self.overlap_classes else main
                if st is not main:
                    if fork is None:
                        fork = torch.cuda.Event()
                        fork.record(main)
                    st.wait_event(fork)
                with torch.cuda.stream(st):
                    _lib.check(L.cymf_als_cg_dev(_lib.ptr(csr[0]), _lib.ptr(csr[1]), _lib.ptr(order[n_long:]),
                                                 n_rows - n_long, _lib.ptr(x_blk), _lib.ptr(yt), None, None, self.dtype,
                                                 K, ld, self.weight, self.cg_tol, self.cg_max_iter, 4, self.stage_rows,
                                                 _lib.ptr(self.queue[1:]), _lib.ptr(self.d_stats), _lib.stream_ptr()))
                    if st is not main:
                        done = torch.cuda.Event()
                        done.record(st)
                        joins.append(done)
            dual = n_rows > n_long and self.dual_max > 0
            if count or dual:
                if self.kernel_events is not None:
                    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    k0.record(main)
                if count and id(csr[0]) in self._ws:
                    rowinfo, cta_ptr, n_ctas = self._ws[id(csr[0])]
                    _lib.check(L.cymf_als_rows_ws_dev(_lib.ptr(rowinfo), _lib.ptr(cta_ptr), n_ctas, _lib.ptr(csr[1]),
                                                      _lib.ptr(x_blk), _lib.ptr(yt), int(yt.shape[0]), self.dtype, K, ld,
                                                      self.weight, self.cg_tol, self.cg_max_iter, _lib.ptr(self.d_stats),
                                                      _lib.ptr(self.d_debug),
                                                      stream))
                elif count:
                    _lib.check(L.cymf_als_rows_tc_dev(_lib.ptr(csr[0]), _lib.ptr(csr[1]), _lib.ptr(order[start:]), count,
                                                      _lib.ptr(x_blk), _lib.ptr(yt), self.dtype, K, ld, self.weight,
                                                      self.cg_tol, self.cg_max_iter, _lib.ptr(self.queue),
                                                      _lib.ptr(self.d_stats), stream))
                if dual:
                    # the short rows, behind the long ones on the same stream: 1 / 2 / 4 rows per 128-slot tile
                    d128, d64, d32 = self._n_dual[id(csr[0])]
                    _lib.check(L.cymf_als_rows_dual_dev(_lib.ptr(csr[0]), _lib.ptr(csr[1]), _lib.ptr(order[n_long:]),
                                                        d128, d64, d32, _lib.ptr(x_blk), _lib.ptr(yt), self.dtype, K, ld,
                                                        self.weight, self.cg_tol, self.cg_max_iter,
                                                        _lib.ptr(self.queue[4:]), _lib.ptr(self.d_stats), stream))
                if self.kernel_events is not None:
                    k1.record(main)
                    # bytes these launches must move: gathered item vectors + their indices, solved rows + their indptr
                    self.kernel_events.append((int(csr[1].numel()) * (K * es + 4) + (n_rows - start) * (K * es + 8), k0, k1))
        else:
            # streaming CG kernel (f64, or A/B runs): 16 warps per row for the longest rows, then 8, then 4; the
            # classes are independent (disjoint rows, own work queues), the two heavy ones go to side streams
            for c, (width, count) in enumerate(zip((16, 8, 4), classes)):
                if count:
                    st = self._side_streams[c] if (c < 2 and self.overlap_classes) else main
                    if st is not main:
                        if fork is None:
                            fork = torch.cuda.Event()
                            fork.record(main)
                        st.wait_event(fork)
                    with torch.cuda.stream(st):
                        _lib.check(L.cymf_als_cg_dev(_lib.ptr(csr[0]), _lib.ptr(csr[1]), _lib.ptr(order[start:]), count,
                                                     _lib.ptr(x_blk), _lib.ptr(yt), None, None, self.dtype, K, ld,
                                                     self.weight, self.cg_tol, self.cg_max_iter, width, self.stage_rows,
                                                     _lib.ptr(self.queue[c:]), _lib.ptr(self.d_stats), _lib.stream_ptr()))
                        if st is not main:
                            done = torch.cuda.Event()
                            done.record(st)
                            joins.append(done)
                start += count
        for done in joins:
            main.wait_event(done)
        self._mark(side + ":row solvers")
        # back to the original coordinates (this rank's block only; the other ranks receive x~ of the NEXT transform)
        _lib.check(L.cymf_rows_times_matrix_dev(_lib.ptr(x_blk), _lib.ptr(x_blk), _lib.ptr(self.Bbwd), self.dtype, R, ld,
                                                _lib.stream_ptr()))
        self._mark(side + ":back transform")
        self._stale[side] = self.dist is not None

    def _half_transformed(self, side):
        other = "item" if side == "user" else "user"
        if self._pub_side != other:                           # first call, after restore(), or two half sweeps of the
            self._publish(other)                              # same side in a row: (re)publish the fixed side
        self._solve(side)
        self._publish(side)
        self._pub_side = side

    def _half_legacy(self, X_full, R, csr, order, Y_full, Ry, classes):
        """solver = "cg" / "pcg": CG on the untransformed systems (G p product inside every iteration); kept for A/B
        comparisons.  Full replicas in the original coordinates, NCCL collectives."""
        L, K, ld = self._L, self.K, self.ld
        stream = _lib.stream_ptr()
        if self.dist:
            y_blk = Y_full[self.rank * Ry:(self.rank + 1) * Ry]
            _lib.check(L.cymf_gram_dev(_lib.ptr(y_blk), self.dtype, Ry, K, ld, self.wd, 0, _lib.ptr(self.ws),
                                       self.ws.numel(), _lib.ptr(self.g64), None, stream))
            self.dist.all_reduce(self.g64)
            _lib.check(L.cymf_gram_finalize_dev(_lib.ptr(self.g64), self.dtype, K, ld, self.wd, _lib.ptr(self.G), stream))
            add_diag = self.wd
        else:
            _lib.check(L.cymf_gram_dev(_lib.ptr(Y_full), self.dtype, Y_full.shape[0], K, ld, self.wd, 1,
                                       _lib.ptr(self.ws), self.ws.numel(), _lib.ptr(self.g64), _lib.ptr(self.G), stream))
            add_diag = 0.0
        x_blk = X_full[self.rank * R:(self.rank + 1) * R]
        ginv = None
        if self.solver == "pcg":                                 # G^-1 as CG preconditioner (f64 Gauss-Jordan, one CTA)
            _lib.check(L.cymf_spd_inverse_dev(_lib.ptr(self.g64), K, ld, add_diag, self.dtype, _lib.ptr(self.Ginv), stream))
            ginv = self.Ginv
        start = 0
        for c, (width, count) in enumerate(zip((16, 8, 4), classes)):
            if count:
                _lib.check(L.cymf_als_cg_dev(_lib.ptr(csr[0]), _lib.ptr(csr[1]), _lib.ptr(order[start:]), count,
                                             _lib.ptr(x_blk), _lib.ptr(Y_full), _lib.ptr(self.G), _lib.ptr(ginv),
                                             self.dtype, K, ld, self.weight, self.cg_tol, self.cg_max_iter, width,
                                             self.stage_rows, _lib.ptr(self.queue[c:]), _lib.ptr(self.d_stats), stream))
            start += count
        if self.dist:
            self.dist.all_gather_into_tensor(X_full, x_blk)

    def user_half(self):
        if self.solver == "transformed":
            return self._half_transformed("user")
        self._half_legacy(self.dW, self.Ru, self.csr_u, self.order_u, self.dH, self.Ri, self.classes_u)

    def item_half(self):
        if self.solver == "transformed":
            return self._half_transformed("item")
        self._half_legacy(self.dH, self.Ri, self.csr_i, self.order_i, self.dW, self.Ru, self.classes_i)

    def _sync_replicas(self):
        """Bring every rank's copy of dW / dH up to date (the half sweeps only exchange the transformed rows).  A
        collective: download(), dense_f64(), residual() and snapshot() call it on all ranks."""
        if self.dist:
            for side, X_full, R in (("user", self.dW, self.Ru), ("item", self.dH, self.Ri)):
                if self._stale[side]:
                    self.dist.all_gather_into_tensor(X_full, X_full[self.rank * R:(self.rank + 1) * R].clone())
                    self._stale[side] = False

    def epoch(self):
        """One epoch = user half sweep + item half sweep (wmf.pyx:111-112).  From the second epoch on the whole
        sequence (about twenty kernels, the cross-rank barriers included) is replayed as ONE CUDA graph."""
        import torch
        with torch.cuda.device(self.dev):
            if (self.use_graph and self.solver == "transformed" and self._pub_side == "item"
                    and self.kernel_events is None and self.trace is None and self.graph_error is None):
                if self._graph is None:
                    self._capture()
                if self._graph is not None:
                    self._graph.replay()
                    self._stale = {"user": self.dist is not None, "item": self.dist is not None}
                    self.epochs_done += 1
                    return
            self.user_half()
            self.item_half()
        self.epochs_done += 1

    def _capture(self):
        import torch
        try:
            torch.cuda.synchronize(self.dev)
            if self.dist:
                self.dist.barrier()
            g = torch.cuda.CUDAGraph()
            l0 = _lib.launch_count()
            with torch.cuda.graph(g):
                self.user_half()
                self.item_half()
            self.graph_launches = _lib.launch_count() - l0      # kernels of libcymf_b200 per replay
            self._graph = g
        except Exception as exc:                                  # noqa: BLE001 - eager launches are always correct
            import warnings
            self.graph_error = repr(exc)
            self._graph = None
            warnings.warn(f"cymf_b200.WMF: CUDA graph capture of the epoch failed, launching eagerly: {self.graph_error}",
                          RuntimeWarning)
            torch.cuda.synchronize(self.dev)

    def download(self, W, H):
        import torch
        self._sync_replicas()
        with torch.cuda.device(self.dev):
            for dev_m, row_slot, out in ((self.dW, self.d_row_slot_u, W), (self.dH, self.d_row_slot_i, H)):
                _lib.download_factor(dev_m.index_select(0, row_slot), self.K, out)   # back to the caller's row order

    def dense_f64(self):
        """(W, H) as dense float64 DEVICE tensors in the caller's row order -- what the on-device evaluator scores."""
        import torch
        out = []
        self._sync_replicas()
        with torch.cuda.device(self.dev):
            for m, slots in ((self.dW, self.slot_u), (self.dH, self.slot_i)):
                t = torch.empty((m.shape[0], self.K), dtype=torch.float64, device=self.dev)
                _lib.check(self._L.cymf_unpack_rows_dev(_lib.ptr(m), _lib.ptr(t), self.dtype, m.shape[0], self.K,
                                                        self.ld, _lib.stream_ptr()))
                key = id(slots)
                if key not in self._unperm:                       # position of every original row in dealt order
                    pos = np.empty(int((slots >= 0).sum()), np.int64)
                    pos[slots[slots >= 0]] = np.flatnonzero(slots >= 0)
                    self._unperm[key] = torch.from_numpy(pos).to(self.dev)
                out.append(t.index_select(0, self._unperm[key]))
        return tuple(out)

    def residual(self, side, sample=24, seed=0):
        """Parity property at sizes no CPU oracle reaches: the largest relative residual, evaluated in float64 from
        the device-resident factors, of the reference's own per-row system (cymf/wmf.pyx:161-168)
            (Y^T Y + wd I + (w - 1) sum_{c in row} y_c y_c^T) x = w sum_{c in row} y_c
        over `sample` random non-empty rows of this rank's block plus its heaviest row; also asserts that rows without
        entries are zero (wmf.pyx:154-156).  Call right after `user_half()` / `item_half()` (side = "user" / "item")."""
        import torch
        self._sync_replicas()
        gen = torch.Generator(device=self.dev)
        gen.manual_seed(int(seed))
        X_full, R, csr, Y_full = ((self.dW, self.Ru, self.csr_u, self.dH) if side == "user" else
                                  (self.dH, self.Ri, self.csr_i, self.dW))
        K = self.K
        with torch.cuda.device(self.dev):
            Y = Y_full[:, :K].double()
            G = Y.T @ Y + self.wd * torch.eye(K, dtype=torch.float64, device=self.dev)
            ip, ix = csr
            lens = ip[1:] - ip[:-1]
            nonempty = torch.nonzero(lens > 0).flatten()
            if nonempty.numel() == 0:
                return 0.0
            pick = nonempty[torch.randint(0, nonempty.numel(), (int(sample),), device=self.dev, generator=gen)]
            pick = torch.cat([pick, torch.argmax(lens).reshape(1)])     # and the heaviest row of the block
            worst = 0.0
            for q in pick.tolist():
                cols = ix[int(ip[q]):int(ip[q + 1])].long()
                Yr = Y[cols]
                A = G + (self.weight - 1.0) * (Yr.T @ Yr)
                b = self.weight * Yr.sum(0)
                x = X_full[self.rank * R + q, :K].double()
                worst = max(worst, float(torch.linalg.norm(A @ x - b) / torch.linalg.norm(b)))
            empty = torch.nonzero(lens == 0).flatten()
            if empty.numel():                                           # wmf.pyx:154-156
                assert not X_full[self.rank * R + empty[:64]].any()
        return worst

    def snapshot(self):
        self._sync_replicas()
        return self.dW.clone(), self.dH.clone()

    def restore(self, snap):
        self.dW.copy_(snap[0])
        self.dH.copy_(snap[1])
        self._stale = {"user": False, "item": False}
        self._pub_side = None                                 # the fixed side must be transformed again

    def stats(self):
        """(CG iterations summed over rows, rows that stopped at cg_max_iter).  Raises if a Gram matrix Y^T Y + wd I
        was not numerically positive definite (weight_decay = 0 with rank-deficient factors): the reference's dgesv
        has no such requirement, the Cholesky change of variables does."""
        info = int(self.d_info.item())
        if info:
            raise _lib.CymfError(f"WMF: Y^T Y + weight_decay I is not positive definite (pivot {info}); use "
                                 "weight_decay > 0 or solver='cg'")
        s = self.d_stats.cpu().numpy()
        if int(s[1]) >> 40:                                    # the warp-specialised solver gave up on a hand-over
            raise _lib.CymfError("WMF: internal hand-over of cymf_als_rows_ws_dev timed out (count, CTA, site, parity, "
                                 f"counters, thread) = {self.d_debug.cpu().tolist()}; set CYMF_ALS_WS=0 and report")
        return int(s[0]), int(s[1])

    @property
    def bytes_per_epoch(self):
        """Algorithmic bytes of one epoch (both half sweeps), SURVEY.md 8(d):
        per half N (K s + 4) + rows (K s + 8) + n K s."""
        s = 4 if self.dtype == _lib.F32 else 8
        K, N, U, I = self.K, self.nnz, self.U, self.I
        return 2 * N * (K * s + 4) + (U + I) * (K * s + 8) + (U + I) * K * s
