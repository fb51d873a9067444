"""`cymf.GloVe` on a B200: same constructor / `fit` / `_fit_glove` signatures and attributes as the reference
class (cymf/glove.pyx:46-177); the prange loop of `_fit_glove` (glove.pyx:149-156) runs as CUDA kernels
(cymf_b200/csrc/glove.cu) reached through the C ABI of include/cymf_b200.h.

Host logic kept in Python as in the reference: the type check (glove.pyx:88-89), the unseeded init
(glove.pyx:91-94, including the `_bias` sized X.shape[0] quirk), the one-time shuffle (glove.pyx:100) and the
final `W = (W + _W) / 2` (glove.pyx:112).  `_fit_glove` is the parity boundary (caller-supplied arrays).

Keyword-only extras: mode "hogwild" (default) | "replay" (serialized f64, reference operation order),
dtype "float32" | "float64", scatter "auto" | "store" | "red".
"""
import ctypes as C

import numpy as np
from scipy import sparse
from sklearn import utils

from . import _lib


class GloVe(object):
    """
    GloVe: Global Vectors for Word Representation, https://nlp.stanford.edu/projects/glove/

    Attributes:
        num_components (int): A dimensionality of latent vector
        learning_rate (double): A learning rate used in AdaGrad
        alpha (double): See the paper.
        x_max (double): See the paper.
        W (np.ndarray[double, ndim=2]): Word vectors
    """

    def __init__(self, num_components=50, learning_rate=0.01, alpha=0.75, x_max=10.0, *,
                 mode="hogwild", dtype="float32", scatter="auto", max_inflight=None, device=None):
        self.num_components = int(num_components)
        self.learning_rate = float(learning_rate)
        self.alpha = float(alpha)
        self.x_max = float(x_max)
        self.W = None
        if mode not in ("hogwild", "replay"):
            raise ValueError("mode must be 'hogwild' or 'replay'")
        if dtype not in _lib.DTYPES:
            raise ValueError("dtype must be 'float32' or 'float64'")
        if scatter not in ("auto", "store", "red"):
            raise ValueError("scatter must be 'auto', 'store' or 'red'")
        self.mode, self.dtype, self.scatter = mode, dtype, scatter
        self.max_inflight = max_inflight
        self.device = device
        self.loss_ = []

    def fit(self, X, num_epochs, num_threads, verbose=False):
        """
        Training GloVe model with Gradient Descent.

        Args:
            X: A word-word cooccurrence matrix.
            num_epochs (int): A number of epochs.
            num_threads (int): accepted for signature compatibility; the GPU grid replaces the thread pool.
            verbose (bool): Whether to show the progress of training.
        """
        if X is None:
            raise ValueError()
        if not isinstance(X, (sparse.lil_matrix, sparse.csr_matrix, sparse.csc_matrix)):
            raise TypeError("X must be a type of scipy.sparse.*_matrix.")

        K = self.num_components
        self.W = np.random.uniform(low=-0.5, high=0.5, size=(X.shape[0], K)) / K
        self.bias = np.random.uniform(low=-0.5, high=0.5, size=(X.shape[0],)) / K
        _W = np.random.uniform(low=-0.5, high=0.5, size=(X.shape[1], K)) / K
        _bias = np.random.uniform(low=-0.5, high=0.5, size=(X.shape[0],)) / K        # glove.pyx:94 (sic)

        coo = X.tocoo()                       # == X.nonzero() + X.data of the reference for canonical CSR / CSC
        keep = coo.data != 0
        central_words, context_words, counts = coo.row[keep], coo.col[keep], coo.data[keep].astype(np.float64)
        self._fit_glove(*utils.shuffle(central_words, context_words, counts), self.W, self.bias, _W, _bias,
                        num_epochs, self.learning_rate, self.x_max, self.alpha, num_threads, verbose)
        self.W = (self.W + _W) / 2.0

    def _fit_glove(self, central_words, context_words, counts, central_W, central_bias, context_W, context_bias,
                   num_epochs, learning_rate, x_max, alpha, num_threads, verbose):
        """Device replacement of `GloVe._fit_glove` (glove.pyx:117-162); the four arrays are updated in place."""
        from tqdm import tqdm
        sess = GloveSession(central_words, context_words, counts, central_W, central_bias, context_W, context_bias,
                            mode=self.mode, dtype=self.dtype, scatter=self.scatter, max_inflight=self.max_inflight,
                            device=self.device)
        self.loss_ = []
        with tqdm(total=num_epochs, leave=True, ncols=100, disable=not verbose) as progress:
            for iteration in range(num_epochs):
                loss = sess.epoch(learning_rate, x_max, alpha, want_loss=verbose)
                if verbose:
                    self.loss_.append(loss)
                    progress.set_description(f"ITER={iteration+1:{len(str(num_epochs))}}, LOSS: {np.round(loss, 4):.4f}")
                progress.update(1)
        sess.download(central_W, central_bias, context_W, context_bias)

    def save_word2vec_format(self, path, index2word):
        """Save the model as gensim.models.KeyedVectors word2vec text format (glove.pyx:164-177)."""
        # The reference loops in Python: f"{index2word[i]} " + " ".join(map(str, self.W[i])) -- minutes for a
        # 400 k x 300 table.  The C writer formats every value exactly as str(np.float64) does (shortest round-trip
        # digits in CPython's repr layout), so the file is byte-identical; host code only, no GPU involved.
        import locale
        W = np.ascontiguousarray(self.W, dtype=np.float64)
        enc = locale.getpreferredencoding(False)                      # what Path.open("w") would encode with
        words = [str(index2word[i]).encode(enc) for i in range(W.shape[0])]
        if any(b"\0" in w for w in words):
            raise ValueError("words must not contain NUL characters")
        arr = (C.c_char_p * len(words))(*words)
        _lib.check(_lib.lib().cymf_word2vec_write_host(str(path).encode(), W.ctypes.data_as(C.c_void_p), W.shape[0],
                                                      W.shape[1], arr))


class GloveSession(object):
    """Device-resident state of one `_fit_glove` call; `epoch()` enqueues one pass of glove.pyx:151-153."""

    def __init__(self, central, context, counts, W, bW, H, bH, *, mode="hogwild", dtype="float32", scatter="auto",
                 max_inflight=None, device=None):
        torch = _lib.require_cuda()
        self._L = _lib.lib()
        self.dev = dev = torch.device(device if device is not None else "cuda")
        for a in (W, bW, H, bH):
            if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous):
                raise ValueError("central_W, central_bias, context_W, context_bias must be C-contiguous float64 arrays")
        self.replay = mode == "replay"
        self.dtype = _lib.F64 if self.replay else _lib.DTYPES[dtype]
        tdt = torch.float64 if self.dtype == _lib.F64 else torch.float32
        self.K = K = W.shape[1]
        self.ld = ld = _lib.ld_for(K)
        self.N = N = int(central.shape[0])
        self.Vw, self.Vh = W.shape[0], H.shape[0]
        if N and (int(np.max(central)) >= min(self.Vw, bW.shape[0]) or int(np.max(context)) >= min(self.Vh, bH.shape[0])):
            raise IndexError("word index out of range of the factor / bias arrays")
        with torch.cuda.device(dev):
            self.d_c = torch.from_numpy(np.ascontiguousarray(central, np.int32)).to(dev, non_blocking=True)
            self.d_x = torch.from_numpy(np.ascontiguousarray(context, np.int32)).to(dev, non_blocking=True)
            self.d_n = torch.from_numpy(np.ascontiguousarray(counts, np.float64)).to(dev).to(tdt)
            self.dW = _lib.upload_factor(W, self.dtype, dev)
            self.dH = _lib.upload_factor(H, self.dtype, dev)
            self.dbW = torch.from_numpy(bW).to(dev).to(tdt)
            self.dbH = torch.from_numpy(bH).to(dev).to(tdt)
            # optimizer.pyx:91-99: every accumulator starts at ONE (pad columns too: they never move)
            self.aW = torch.ones((self.Vw, ld), dtype=tdt, device=dev)
            self.aH = torch.ones((self.Vh, ld), dtype=tdt, device=dev)
            self.abW = torch.ones(bW.shape[0], dtype=tdt, device=dev)
            self.abH = torch.ones(bH.shape[0], dtype=tdt, device=dev)
            self.d_loss = torch.zeros(1, dtype=torch.float64, device=dev)
        self.p = _lib.GloveParams(*[_lib.ptr(t) for t in (self.dW, self.dH, self.dbW, self.dbH, self.aW, self.aH,
                                                             self.abW, self.abH)])
        self.scatter = {"auto": 1 if self.dtype == _lib.F32 else 0, "store": 0, "red": 1}[scatter]
        self.inflight = int(max_inflight) if max_inflight is not None else max(1024, N // 256)
        self.epochs_done = 0

    def epoch(self, learning_rate, x_max, alpha, want_loss=False):
        import torch
        L = self._L
        with torch.cuda.device(self.dev):
            stream = _lib.stream_ptr()
            if self.replay:
                loss = torch.empty(self.N, dtype=torch.float64, device=self.dev) if want_loss else None
                _lib.check(L.cymf_glove_replay_epoch_dev(C.byref(self.p), _lib.ptr(self.d_c), _lib.ptr(self.d_x),
                                                         _lib.ptr(self.d_n), self.N, self.K, self.ld, learning_rate,
                                                         x_max, alpha, _lib.ptr(loss), stream))
                out = float(loss.sum().item()) / max(self.N, 1) if want_loss else None
            else:
                if want_loss:
                    self.d_loss.zero_()
                _lib.check(L.cymf_glove_hogwild_epoch_dev(C.byref(self.p), self.dtype, self.scatter, _lib.ptr(self.d_c),
                                                          _lib.ptr(self.d_x), _lib.ptr(self.d_n), self.N, self.K,
                                                          self.ld, learning_rate, x_max, alpha, self.inflight,
                                                          _lib.ptr(self.d_loss) if want_loss else None, stream))
                out = float(self.d_loss.item()) / max(self.N, 1) if want_loss else None
        self.epochs_done += 1
        return out

    def download(self, W, bW, H, bH):
        import torch
        with torch.cuda.device(self.dev):
            _lib.download_factor(self.dW, self.K, W)
            _lib.download_factor(self.dH, self.K, H)
            bW[...] = self.dbW.to(torch.float64).cpu().numpy()
            bH[...] = self.dbH.to(torch.float64).cpu().numpy()

    @property
    def bytes_per_sample(self):
        """Algorithmic bytes per sample: 4 rows RW + 4 scalars RW + (c, x, count)."""
        s = 4 if self.dtype == _lib.F32 else 8
        return 8 * self.K * s + 8 * s + 8 + s


def _vocabulary_pass(raw, min_count):
    """The Python part of the reference's `read_text` (cymf/glove.pyx:198-214) with array operations instead of a
    per-token interpreter loop (21 s -> ~3 s for a text8-sized corpus), same results in every case:

        count  = Counter(raw.replace("\\n", "<eos>").split(" "))        # newlines GLUE the neighbouring words
        for line in raw.split("\\n"): for w in line.split(" "):
            count[w]                      -> KeyError for a word that only ever occurs next to a newline
            kept iff count[w] >= min_count; ids in first-appearance order of the kept words

    Returns (tokens int32[T] = ids of the kept words, lines concatenated; pos int32[T] = index of each kept word
    inside its line; i2w dict in id order)."""
    import pandas as pd
    glued = raw.replace("\n", "<eos>").split(" ")
    lines = raw.split("\n")
    codes_g, uniq_g = pd.factorize(np.array(glued, dtype=object))     # codes in first-appearance order
    if len(lines) == 1:
        codes, uniq, lens = codes_g, uniq_g, np.array([len(glued)], np.int64)
    else:
        per_line = [line.split(" ") for line in lines]
        lens = np.fromiter((len(w) for w in per_line), np.int64, len(per_line))
        flat = [w for ws in per_line for w in ws]
        codes, uniq = pd.factorize(np.array(flat, dtype=object))
    count = dict(zip(uniq_g.tolist(), np.bincount(codes_g, minlength=len(uniq_g)).tolist()))
    cnt_u = np.fromiter((count.get(w, -1) for w in uniq.tolist()), np.int64, len(uniq))
    if (cnt_u < 0).any():                                             # first word, in corpus order, without a count
        first = int(np.flatnonzero(cnt_u[codes] < 0)[0])
        raise KeyError(uniq[codes[first]])
    keep_u = cnt_u >= min_count
    new_id = np.cumsum(keep_u) - 1                                    # kept words keep their order of first appearance
    keep = keep_u[codes]
    tokens = new_id[codes[keep]].astype(np.int32)
    kept_before = np.cumsum(keep) - keep                              # kept words before each word, corpus-wide
    line_first = np.concatenate([[0], np.cumsum(lens)[:-1]])          # every line has >= 1 word ("".split(" ") == [""])
    line_of = np.repeat(np.arange(lens.shape[0]), lens)
    pos = (kept_before - kept_before[line_first][line_of])[keep].astype(np.int32)
    words = uniq.tolist()
    i2w = {int(new_id[k]): words[k] for k in np.flatnonzero(keep_u).tolist()}
    return tokens, pos, i2w


def read_text(fname, min_count=5, window_size=10):
    """`cymf.glove.read_text(fname, min_count, window_size)` (cymf/glove.pyx:183-241): co-occurrence matrix of a
    text file -> (scipy.sparse.csr_matrix [V, V] float64, i2w dict).

    The vocabulary pass keeps the reference's semantics (word counts over the text with newlines glued as "<eos>",
    ids in first-appearance order of the words with count >= min_count, KeyError for a word that only ever occurs
    next to a newline, glove.pyx:198-214) with array operations (`_vocabulary_pass`; checked against the
    statement-for-statement restatement in oracle/oracle.py).  The counting loop (glove.pyx:218-221, an
    `unordered_map<long, double>` updated once per (token, earlier token within the window): 170 M updates for text8)
    runs on the device: cymf_cooc_count_dev sorts the updates by cell with two stable radix sorts and sums every
    cell in corpus order in f64, so the counts are bit-identical to the reference's."""
    import torch
    with open(fname) as f:
        raw = f.read()
    tokens, pos, i2w = _vocabulary_pass(raw, min_count)
    w2i = i2w
    V = len(w2i)
    T = len(tokens)
    if V == 0 or T == 0:
        return sparse.csr_matrix((V, V), dtype=np.float64), i2w
    _lib.require_cuda()
    L = _lib.lib()
    dev = torch.device("cuda", torch.cuda.current_device())
    d_tok = torch.from_numpy(np.ascontiguousarray(tokens, np.int32)).to(dev)
    d_pos = torch.from_numpy(np.ascontiguousarray(pos, np.int32)).to(dev)
    cap = T * int(window_size)
    rows = torch.empty(cap, dtype=torch.int32, device=dev)
    cols = torch.empty(cap, dtype=torch.int32, device=dev)
    vals = torch.empty(cap, dtype=torch.float64, device=dev)
    nnz = torch.zeros(1, dtype=torch.int64, device=dev)
    ws = torch.empty(max(int(L.cymf_cooc_workspace_bytes(T, int(window_size))), 256), dtype=torch.uint8, device=dev)
    _lib.check(L.cymf_cooc_count_dev(_lib.ptr(d_tok), _lib.ptr(d_pos), T, V, int(window_size), _lib.ptr(rows), _lib.ptr(cols),
                                     _lib.ptr(vals), cap, _lib.ptr(nnz), _lib.ptr(ws), _lib.stream_ptr()))
    m = int(nnz.item())
    r, c, v = rows[:m].cpu().numpy(), cols[:m].cpu().numpy(), vals[:m].cpu().numpy()
    indptr = np.concatenate([[0], np.cumsum(np.bincount(r, minlength=V))]).astype(np.int32 if m < 2 ** 31 else np.int64)
    X = sparse.csr_matrix((v, c, indptr), shape=(V, V))       # cells arrive sorted by (row, col): canonical CSR
    X.has_sorted_indices = True
    return X, i2w
