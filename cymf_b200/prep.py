"""Sparse-matrix preparation on the device (cymf_b200/csrc/prep.cu): CSR transpose (`X.T.tocsr()`, which the
reference evaluates twice per epoch on the host, cymf/wmf.pyx:112), the heaviest-first row deal of the sharded ALS
and the relabelled row blocks.  Inputs and outputs are torch DEVICE tensors (indptr int64, indices int32); results
are exact, deterministic and equal to the host (scipy / NumPy) constructions they replace -- tests/test_prep_gpu.py.
"""
import torch

from . import _lib


def _workspace(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def exclusive_scan_u32(counts):
    """int64 [n+1] exclusive prefix sums (last entry = total) of a uint32-valued int32/uint32 device tensor."""
    L = _lib.lib()
    n = counts.numel()
    out = torch.empty(n + 1, dtype=torch.int64, device=counts.device)
    ws = _workspace(L.cymf_scan_workspace_bytes(n), counts.device)
    _lib.check(L.cymf_exclusive_scan_u32_dev(_lib.ptr(counts), _lib.ptr(out), n, _lib.ptr(ws), _lib.stream_ptr()))
    return out


def sort_pairs(keys, values, key_bits=32):
    """Stable in-place sort of (keys, values) (int32 tensors read as uint32) by the low `key_bits` bits of the key."""
    L = _lib.lib()
    n = keys.numel()
    ws = _workspace(L.cymf_sort_workspace_bytes(n), keys.device)
    _lib.check(L.cymf_sort_pairs_dev(_lib.ptr(keys), _lib.ptr(values), n, int(key_bits), _lib.ptr(ws), _lib.stream_ptr()))
    return keys, values


def transpose_csr(indptr, indices, rows, cols):
    """(t_indptr int64[cols+1], t_indices int32[nnz]) of X^T, rows sorted: scipy's `X.T.tocsr()`."""
    L = _lib.lib()
    nnz = indices.numel()
    dev = indptr.device
    t_indptr = torch.empty(cols + 1, dtype=torch.int64, device=dev)
    t_indices = torch.empty(nnz, dtype=torch.int32, device=dev)
    ws = _workspace(L.cymf_csr_transpose_workspace_bytes(rows, cols, nnz), dev)
    _lib.check(L.cymf_csr_transpose_dev(_lib.ptr(indptr), _lib.ptr(indices), rows, cols, nnz, _lib.ptr(t_indptr),
                                        _lib.ptr(t_indices), _lib.ptr(ws), _lib.stream_ptr()))
    return t_indptr, t_indices


def deal_rows(indptr, rows, world):
    """Heaviest-first round-robin deal.  Returns (slot_row int64[world*R], row_slot int64[rows], R)."""
    L = _lib.lib()
    dev = indptr.device
    R = (rows + world - 1) // world
    slot_row = torch.empty(R * world, dtype=torch.int64, device=dev)
    row_slot = torch.empty(max(rows, 1), dtype=torch.int64, device=dev)[:rows]
    ws = _workspace(L.cymf_deal_workspace_bytes(rows), dev)
    _lib.check(L.cymf_deal_rows_dev(_lib.ptr(indptr), rows, world, R, _lib.ptr(slot_row), _lib.ptr(row_slot),
                                    _lib.ptr(ws), _lib.stream_ptr()))
    return slot_row, row_slot, R


def csr_block(indptr, indices, slot_row, col_slot=None):
    """CSR (int64 indptr, int32 indices) whose row q is source row slot_row[q] (-1 = empty) with columns renamed
    through col_slot."""
    L = _lib.lib()
    dev = indptr.device
    slot_row = slot_row.contiguous()
    count = slot_row.numel()
    blk_indptr = torch.empty(count + 1, dtype=torch.int64, device=dev)
    ws = _workspace(L.cymf_csr_block_workspace_bytes(count), dev)
    args = (_lib.ptr(indptr), _lib.ptr(indices), _lib.ptr(slot_row), count, _lib.ptr(col_slot), _lib.ptr(blk_indptr))
    _lib.check(L.cymf_csr_block_dev(*args, None, _lib.ptr(ws), _lib.stream_ptr()))
    nnz = int(blk_indptr[-1].item())
    blk_indices = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)[:nnz]
    if nnz:
        _lib.check(L.cymf_csr_block_dev(*args, _lib.ptr(blk_indices), _lib.ptr(ws), _lib.stream_ptr()))
    return blk_indptr, blk_indices
