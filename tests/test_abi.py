"""CPU-side checks of the drop-in boundary: the shared library loads, exports every symbol that
include/cymf_b200.h declares, the Python binding table covers all of them, and the host-side pieces
(mt19937 negative stream, argument validation, loud failure without a GPU) behave."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden
from cymf_b200 import _lib

HEADER = os.path.join(ROOT, "include", "cymf_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cymf_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) >= 10
    handle = ctypes.CDLL(_lib.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), f"{n} declared in include/cymf_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "cymf_b200/_lib.py SIGNATURES out of sync with the header"
    assert _lib.lib().cymf_abi_version() == 1


def test_host_rng_matches_libstdcxx_vectors():
    g = golden("rng.npz")
    for n in (1682, 26744, 7, 1000003):
        assert np.array_equal(_lib.HostRng(1234).below(n, 64), g[f"below_{n}_seed1234"])
    assert np.array_equal(_lib.HostRng(99).below(3, 200), g["below_3_seed99"])
    r = _lib.HostRng(1234)            # the stream persists across calls (one generator per fit, bpr.pyx:141)
    a, b = r.below(1682, 10), r.below(1682, 54)
    assert np.array_equal(np.concatenate([a, b]), g["below_1682_seed1234"])


def test_host_rng_long_stream_matches_oracle(oracle):
    assert np.array_equal(_lib.HostRng(4242).below(26744, 200_000), oracle.Rng(4242).below(26744, 200_000))


def test_argument_errors_are_reported_not_fatal():
    L = _lib.lib()
    assert L.cymf_rng_fill_below(None, 5, None, 1) == -1
    assert b"bad argument" in L.cymf_last_error()
    f = _lib.Factors()
    rc = L.cymf_bpr_hogwild_epoch_dev(ctypes.byref(f), 0, 0, 0, None, None, 0, None, None, 1, 1, 1, 4, 0.1, 0.1, 0, 0,
                                      0, None, None)
    assert rc == -1 and b"null pointer" in L.cymf_last_error()


def test_constructor_contract():
    import cymf_b200 as cymf
    with pytest.raises(Exception, match="rmsprop is invalid."):
        cymf.BPR(optimizer="rmsprop")
    m = cymf.BPR()
    assert (m.num_components, m.learning_rate, m.optimizer, m.weight_decay) == (20, 0.001, "adam", 0.01)
    assert m.W is None and m.H is None
    with pytest.raises(ValueError):
        m.fit(None)
    with pytest.raises(ValueError):
        m.fit([[1, 0], [0, 1]])
    with pytest.raises(ValueError):
        m.fit(np.eye(3), early_stopping=True)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import cymf_b200 as cymf
    with pytest.raises(_lib.CymfError, match="no CPU fallback"):
        cymf.BPR(4, optimizer="sgd").fit(np.eye(4), num_epochs=1, verbose=False)


def test_host_rng_wide_ranges_match_libstdcxx_vectors():
    """cymf_rng_fill_below64: the raw-word (n = 2^32) and upscaling (n > 2^32) branches of libstdc++'s
    uniform_int_distribution<long>, which RelMF's replay needs for U*I cells (cymf/relmf.pyx:127)."""
    g = golden("rng64.npz")
    for n in (3703857792, 4294967296, 4294967297, 5000000000, 10000000000000):
        want = g[f"below_{n}_seed1234"]
        assert np.array_equal(_lib.HostRng(1234).below64(n, want.shape[0]), want), n
    r = _lib.HostRng(1234)                      # narrow ranges go through the same 32-bit path as below()
    assert np.array_equal(r.below64(1682, 64), golden("rng.npz")["below_1682_seed1234"])


def test_relmf_contract_and_propensities():
    import cymf_b200 as cymf
    from cymf_b200.relmf import item_propensities
    from scipy import sparse
    with pytest.raises(Exception, match="rmsprop is invalid."):
        cymf.RelMF(optimizer="rmsprop")                                       # relmf.pyx:65-66
    m = cymf.RelMF()
    assert (m.num_components, m.clip_value, m.learning_rate, m.optimizer, m.weight_decay) == (20, 0.1, 0.001, "adam", 0.01)
    assert m.W is None and m.H is None
    with pytest.raises(ValueError):
        m.fit(None)
    rng = np.random.default_rng(3)
    A = (rng.random((50, 40)) < 0.2) * rng.integers(1, 6, (50, 40)) / 5.0
    A[:, 7] = 0                                                                # an item nobody has: floor 1e-5
    want = np.maximum(A.mean(axis=0) / A.mean(axis=0).max(), 1e-5) ** 0.5     # relmf.pyx:90 on the dense matrix
    assert np.array_equal(item_propensities(sparse.csr_matrix(A)), want)


def test_read_text_without_vocabulary_needs_no_gpu(tmp_path):
    """min_count above every word count: the reference returns an empty 0 x 0 matrix and an empty map."""
    import cymf_b200 as cymf
    f = tmp_path / "tiny.txt"
    f.write_text("a b c a b")
    X, i2w = cymf.glove.read_text(str(f), 5, 3)
    assert X.shape == (0, 0) and X.nnz == 0 and i2w == {}
    with pytest.raises(KeyError):                                              # glove.pyx:199-209: "c<eos>d" hides c and d
        g = tmp_path / "two_lines.txt"
        g.write_text("a b c\nd a b")
        cymf.glove.read_text(str(g), 1, 2)


def test_wmf_and_relmf_argument_validation():
    import cymf_b200 as cymf
    with pytest.raises(ValueError):
        cymf.WMF(prep="somewhere")
    with pytest.raises(ValueError):
        cymf.WMF(257)
    with pytest.raises(ValueError):
        cymf.RelMF(mode="serial")
    L = _lib.lib()
    assert L.cymf_sort_pairs_dev(None, None, 5, 8, None, None) == -1 and b"bad argument" in L.cymf_last_error()
    assert L.cymf_cooc_count_dev(None, None, 1 << 30, 10, 8, None, None, None, 0, None, None, None) == -1
    assert L.cymf_als_heavy_workspace_doubles(3, 128, 128) == 3 * (128 * 128 + 128)
    assert L.cymf_cooc_workspace_bytes(1000, 10) > 10000 * 20


def _oracle_tokens(oracle, path, min_count):
    x, i2w = oracle.read_text_vocabulary(path, min_count)
    tokens = np.array([t for line in x for t in line], np.int32)
    pos = np.array([p for line in x for p in range(len(line))], np.int32)
    return tokens, pos, i2w


@pytest.mark.parametrize("seed", range(12))
def test_vocabulary_pass_equals_statement_for_statement_restatement(oracle, tmp_path, seed):
    """cymf_b200.glove._vocabulary_pass (array operations) against the reference's loop restated statement for
    statement (oracle.read_text_vocabulary, cymf/glove.pyx:198-214): random corpora with several lines, empty lines,
    double spaces, rare words and words glued by newlines -- same ids, same positions, same KeyError."""
    from cymf_b200.glove import _vocabulary_pass
    rng = np.random.default_rng(seed)
    vocab = [f"w{i}" for i in range(int(rng.integers(3, 40)))] + [""] * int(rng.integers(0, 3))
    n_lines = int(rng.integers(1, 6))
    lines = []
    for _ in range(n_lines):
        n = int(rng.integers(0, 60))
        p = 1.0 / (np.arange(len(vocab)) + 1.0)
        lines.append(" ".join(rng.choice(vocab, size=n, p=p / p.sum()).tolist()))
    raw = "\n".join(lines) + ("\n" if rng.random() < 0.3 else "")
    f = tmp_path / "c.txt"
    f.write_text(raw)
    min_count = int(rng.integers(1, 4))
    try:
        want = _oracle_tokens(oracle, str(f), min_count)
    except KeyError as e:
        with pytest.raises(KeyError) as got:
            _vocabulary_pass(raw, min_count)
        assert got.value.args == e.args
        return
    tokens, pos, i2w = _vocabulary_pass(raw, min_count)
    assert np.array_equal(tokens, want[0]) and np.array_equal(pos, want[1])
    assert i2w == want[2] and list(i2w) == list(want[2])                       # same map, same insertion order


def test_vocabulary_pass_on_the_golden_corpora(oracle):
    from conftest import GOLDEN
    from cymf_b200.glove import _vocabulary_pass
    for txt, mc in (("corpus_one_line.txt", 5), ("corpus_lines.txt", 3)):
        path = os.path.join(GOLDEN, txt)
        tokens, pos, i2w = _vocabulary_pass(open(path).read(), mc)
        want = _oracle_tokens(oracle, path, mc)
        assert np.array_equal(tokens, want[0]) and np.array_equal(pos, want[1]) and i2w == want[2]


def test_save_word2vec_format_is_byte_identical(tmp_path):
    """GloVe.save_word2vec_format through the C writer == the reference's Python loop (cymf/glove.pyx:164-177), byte for
    byte: str(np.float64) is the shortest round-trip decimal in CPython's repr layout; random bit patterns over the
    whole double range, layout thresholds (1e-4 / 1e-5, 1e15 / 1e16 / 1e17), zeros, subnormals, inf, nan."""
    import cymf_b200 as cymf
    rng = np.random.default_rng(0)
    n = 60_000
    bits = rng.integers(0, 2 ** 63, n, dtype=np.uint64) | (rng.integers(0, 2, n, dtype=np.uint64) << np.uint64(63))
    special = np.array([0.0, -0.0, 1e-5, 9.999e-5, 1e-4, 0.001, 1e15, 1e16, 9999999999999998.0, 1e17, np.inf, -np.inf,
                        np.nan, 5e-324, 1.7976931348623157e308, 2.2250738585072014e-308, 1.0, -1.0, 0.1, 123456.789,
                        1e22, 1e23, 100.0, 12345678901234567890.0, 0.30000000000000004])
    vals = np.concatenate([bits.view(np.float64), rng.normal(size=30_000), rng.normal(size=10_000) * 1e-5,
                           10.0 ** rng.integers(-25, 25, 5_000).astype(float), special])
    K = 25
    W = vals[:(vals.shape[0] // K) * K].reshape(-1, K)
    i2w = {i: f"wörd{i}" for i in range(W.shape[0])}
    ref = tmp_path / "ref.vec"
    with ref.open("w") as f:                                                   # the reference's loop, verbatim
        f.write(f"{W.shape[0]} {W.shape[1]}\n")
        for i in range(W.shape[0]):
            f.write(f"{i2w[i]} " + " ".join(list(map(str, W[i]))) + "\n")
    m = cymf.GloVe(K)
    m.W = W
    out = tmp_path / "new.vec"
    m.save_word2vec_format(str(out), i2w)
    assert out.read_bytes() == ref.read_bytes()
    with pytest.raises(_lib.CymfError):
        m.save_word2vec_format(str(tmp_path / "missing_dir" / "x.vec"), i2w)


def test_synthetic_movielens_has_the_reference_dataset_shape():
    import cymf_b200 as cymf
    d = cymf.dataset.SyntheticMovieLens("ml-100k")
    assert d.train.shape == d.valid.shape == d.test.shape == (943, 1682) == (d.num_user, d.num_item)
    assert d.train.format == "lil" and (d.train_size, d.valid_size, d.test_size) == (d.train.nnz, d.valid.nnz, d.test.nnz)
    total = d.train_size + d.valid_size + d.test_size
    assert abs(d.test_size / total - 0.1) < 0.01 and abs(d.valid_size / total - 0.09) < 0.01
    assert d.train.multiply(d.test).nnz == 0 and d.train.multiply(d.valid).nnz == 0      # disjoint splits
    with pytest.raises(ValueError):
        cymf.dataset.SyntheticMovieLens("ml-9000")
    with pytest.raises(RuntimeError):
        cymf.dataset.MovieLens("ml-100k")


@pytest.mark.parametrize("n,n_ctas,seed", [(0, 4, 0), (1, 4, 1), (5, 8, 2), (1000, 7, 3), (20000, 148, 4)])
def test_ws_row_schedule_is_a_balanced_permutation(n, n_ctas, seed):
    """cymf_als_ws_schedule_host (host-only): every row lands in exactly one CTA's list with its extent, the loads
    (entries + row_cost per row) differ by about one row, a row far above the mean gets a CTA (almost) to itself, and
    inside a list long and short rows alternate (longest, shortest, second longest, ...)."""
    import ctypes as C
    from cymf_b200 import _lib
    L = _lib.lib()
    rng = np.random.default_rng(seed)
    lens = np.clip(np.rint(rng.lognormal(0.0, 1.2, n) * 150), 0, None).astype(np.int64)
    if n >= 1000:
        lens[rng.integers(0, n)] = 40 * int(lens.sum() // n_ctas // 10 + 1)          # one very long row (~4x the mean load)
    indptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    rows = rng.permutation(n).astype(np.int32)                                       # any row order
    cta_ptr = np.full(n_ctas + 1, -1, np.int32)
    rowinfo = np.full(max(4 * n, 1), -1, np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    assert L.cymf_als_ws_schedule_host(p(indptr), p(rows), n, n_ctas, 256, p(cta_ptr), p(rowinfo)) == 0
    assert cta_ptr[0] == 0 and cta_ptr[-1] == n and (np.diff(cta_ptr) >= 0).all()
    ri = rowinfo[:4 * n].reshape(n, 4)
    assert np.array_equal(np.sort(ri[:, 0]), np.arange(n))                            # a permutation of the rows
    assert np.array_equal(ri[:, 1], lens[ri[:, 0]])                                   # nnz
    lo = ri[:, 2].astype(np.uint32).astype(np.int64) | (ri[:, 3].astype(np.int64) << 32)
    assert np.array_equal(lo, indptr[ri[:, 0]])                                       # 64-bit row start, two words
    if n == 0:
        return
    loads = np.array([int((ri[cta_ptr[b]:cta_ptr[b + 1], 1] + 256).sum()) for b in range(n_ctas)])
    mean = loads.sum() / n_ctas
    big = int(lens.max()) + 256
    if n < n_ctas:                                                                   # fewer rows than CTAs: one row each
        assert (np.diff(cta_ptr) <= 1).all()
    elif big <= mean:
        assert loads.max() - loads.min() <= 2 * big
    else:                                                                            # the long row's CTA holds (nearly) only it
        assert loads.max() <= big + 2 * (int(np.sort(lens)[-2]) + 256)
        others = np.delete(loads, loads.argmax())
        assert others.max() - others.min() <= 2 * (int(np.sort(lens)[-2]) + 256)
    for b in range(n_ctas):                                                           # longest, shortest, 2nd longest, ...
        seg = ri[cta_ptr[b]:cta_ptr[b + 1], 1]
        if seg.shape[0] >= 4:
            assert (np.diff(seg[0::2]) <= 0).all() and (np.diff(seg[1::2]) >= 0).all()
            assert seg[0] == seg.max() and seg[1] == seg.min()
    with pytest.raises(_lib.CymfError):
        _lib.check(L.cymf_als_ws_schedule_host(p(indptr), p(rows), n, 0, 256, p(cta_ptr), p(rowinfo)))
