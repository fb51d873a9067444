"""Device-side sparse preparation (cymf_b200/csrc/prep.cu) against the host constructions it replaces: scipy's
`X.T.tocsr()` (cymf/wmf.pyx:112), NumPy's stable argsort deal and the relabelled row blocks (pytest -m gpu).
Everything is integer work: the bar is exact equality."""
import numpy as np
import pytest
from scipy import sparse

pytestmark = pytest.mark.gpu


def _dev(a, dt):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a, dt)).cuda()


@pytest.mark.parametrize("n", [0, 1, 5, 2047, 2048, 2049, 100_000, 3_000_001])
def test_exclusive_scan(n):
    from cymf_b200 import prep
    rng = np.random.default_rng(n)
    x = rng.integers(0, 1 << 20, n).astype(np.uint32)
    got = prep.exclusive_scan_u32(_dev(x.view(np.int32), np.int32)).cpu().numpy()
    want = np.concatenate([[0], np.cumsum(x.astype(np.int64))])
    assert np.array_equal(got, want)


@pytest.mark.parametrize("n,bits", [(1, 8), (33, 8), (4096, 16), (4097, 20), (1_000_003, 24), (2_500_000, 32)])
def test_radix_sort_pairs_is_stable(n, bits):
    from cymf_b200 import prep
    rng = np.random.default_rng(n)
    keys = (rng.integers(0, 1 << bits, n, dtype=np.uint64) >> (rng.integers(0, bits, n).astype(np.uint64))).astype(np.uint32)
    vals = np.arange(n, dtype=np.uint32)
    k, v = prep.sort_pairs(_dev(keys.view(np.int32), np.int32), _dev(vals.view(np.int32), np.int32), bits)
    order = np.argsort(keys, kind="stable")
    assert np.array_equal(k.cpu().numpy().view(np.uint32), keys[order])
    assert np.array_equal(v.cpu().numpy().view(np.uint32), vals[order])        # equal keys keep their input order


@pytest.mark.parametrize("shape", [(1, 1, 1), (7, 5, 0), (60, 90, 700), (943, 1682, 100_000), (6040, 3706, 1_000_000)])
def test_transpose_equals_scipy(shape):
    import cymf_b200 as cymf
    from cymf_b200 import prep
    U, I, nnz = shape
    X = cymf.synth.synth_implicit(U, I, nnz, seed=9) if nnz else sparse.csr_matrix((U, I))
    if nnz > 100:
        X = X.tolil(); X[U // 2, :] = 0; X[:, I // 3] = 0; X = X.tocsr(); X.eliminate_zeros()      # empty row / column
    XT = X.T.tocsr()
    XT.sort_indices()
    t_ip, t_ix = prep.transpose_csr(_dev(X.indptr, np.int64), _dev(X.indices, np.int32), U, I)
    assert np.array_equal(t_ip.cpu().numpy(), XT.indptr.astype(np.int64))
    assert np.array_equal(t_ix.cpu().numpy(), XT.indices.astype(np.int32))
    # transposing twice gives X back (size-independent property, also checked at full size below)
    b_ip, b_ix = prep.transpose_csr(t_ip, t_ix, I, U)
    assert np.array_equal(b_ip.cpu().numpy(), X.indptr) and np.array_equal(b_ix.cpu().numpy(), X.indices)


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_deal_and_blocks_equal_host_construction(world):
    import cymf_b200 as cymf
    from cymf_b200 import prep
    from cymf_b200.wmf import _deal, _relabel
    X = cymf.synth.synth_implicit(301, 157, 6000, seed=4).tolil()
    X[5, :] = 0
    X = X.tocsr(); X.eliminate_zeros()
    XT = X.T.tocsr(); XT.sort_indices()
    U, I = X.shape
    slot_u, Ru = _deal(np.diff(X.indptr), world)
    slot_i, Ri = _deal(np.diff(XT.indptr), world)
    new_i = np.empty(I, np.int64); new_i[slot_i[slot_i >= 0]] = np.flatnonzero(slot_i >= 0)
    ip, ix = _dev(X.indptr, np.int64), _dev(X.indices, np.int32)
    d_slot_u, d_row_slot_u, dRu = prep.deal_rows(ip, U, world)
    t_ip, t_ix = prep.transpose_csr(ip, ix, U, I)
    d_slot_i, d_row_slot_i, dRi = prep.deal_rows(t_ip, I, world)
    assert (dRu, dRi) == (Ru, Ri)
    assert np.array_equal(d_slot_u.cpu().numpy(), slot_u) and np.array_equal(d_slot_i.cpu().numpy(), slot_i)
    assert np.array_equal(d_row_slot_i.cpu().numpy(), new_i)
    for rank in range(world):
        want = _relabel(X, slot_u[rank * Ru:(rank + 1) * Ru], new_i, slot_i.shape[0])
        b_ip, b_ix = prep.csr_block(ip, ix, d_slot_u[rank * Ru:(rank + 1) * Ru], d_row_slot_i)
        assert np.array_equal(b_ip.cpu().numpy(), want.indptr.astype(np.int64))
        assert np.array_equal(b_ix.cpu().numpy(), want.indices.astype(np.int32))


def test_wmf_device_prep_equals_host_prep(oracle):
    """Same factors whichever side builds X^T / the deal / the blocks (row order inside a block row is identical,
    so even f32 sums agree bit for bit)."""
    import cymf_b200 as cymf
    X = cymf.synth.synth_implicit(400, 300, 9000, seed=6)
    out = []
    for prep_side in ("host", "device"):
        m = cymf.WMF(32, 0.01, 10.0, prep=prep_side)
        m.fit(X, 2, 1, verbose=False)
        out.append((m.W.copy(), m.H.copy()))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    Wo, Ho = oracle.wmf_fit(X, 32, 0.01, 10.0, 2)
    assert np.abs(out[1][0] - Wo).max() <= 1e-4 * np.abs(Wo).max()


def test_transpose_involution_at_c3_size():
    """ml-20m shape (20 M nonzeros), data built on the device: (X^T)^T == X, column counts match, rows sorted."""
    import torch
    from cymf_b200 import prep
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    U, I, per = 138_493, 26_744, 144
    cols = torch.randint(0, I, (U, per), device="cuda", generator=g, dtype=torch.int32)
    cols = torch.sort(cols, dim=1).values
    keep = torch.ones_like(cols, dtype=torch.bool); keep[:, 1:] = cols[:, 1:] != cols[:, :-1]       # dedupe per row
    ip = torch.cat([torch.zeros(1, dtype=torch.int64, device="cuda"), keep.sum(1).cumsum(0)])
    ix = cols[keep].contiguous()
    t_ip, t_ix = prep.transpose_csr(ip, ix, U, I)
    assert int(t_ip[-1]) == ix.numel()
    assert torch.equal(t_ip[1:] - t_ip[:-1], torch.bincount(ix.long(), minlength=I))
    b_ip, b_ix = prep.transpose_csr(t_ip, t_ix, I, U)
    assert torch.equal(b_ip, ip) and torch.equal(b_ix, ix)
