"""ctypes binding of cymf_b200/libcymf_b200.so (the C ABI of include/cymf_b200.h).

There is no fallback of any kind: if the library is missing, or there is no CUDA device, every compute
call raises.  Build the library with `python -c "import __graft_entry__ as g; g.build()"` or
`make -C cymf_b200/csrc`.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcymf_b200.so")

F32, F64 = 0, 1
SGD, ADAGRAD, ADAM = 0, 1, 2
OPTIMIZERS = {"sgd": SGD, "adagrad": ADAGRAD, "adam": ADAM}
DTYPES = {"float32": F32, "float64": F64}

_p, _i32, _i64, _u32, _u64, _f64 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_uint64, C.c_double


class Factors(C.Structure):
    """struct cymf_factors"""
    _fields_ = [(n, _p) for n in ("W", "H", "s1W", "s1H", "s2W", "s2H")]


class GloveParams(C.Structure):
    """struct cymf_glove_params"""
    _fields_ = [(n, _p) for n in ("W", "H", "bW", "bH", "aW", "aH", "abW", "abH")]


# name -> (restype, argtypes); must list EVERY symbol include/cymf_b200.h declares (tests check this)
SIGNATURES = {
    "cymf_abi_version": (C.c_int, []),
    "cymf_last_error": (C.c_char_p, []),
    "cymf_launch_count": (_i64, []),
    "cymf_rng_create": (_p, [_u32]),
    "cymf_rng_destroy": (None, [_p]),
    "cymf_rng_fill_below": (C.c_int, [_p, _u32, _p, _i64]),
    "cymf_rng_fill_below64": (C.c_int, [_p, _u64, _p, _i64]),
    "cymf_convert_dev": (C.c_int, [_p, _p, C.c_int, _i64, _p]),
    "cymf_relmf_hogwild_epoch_dev": (C.c_int, [C.POINTER(Factors), C.c_int, C.c_int, C.c_int, _p, _p, _p, _p,
                                               _i32, _i32, _i32, _i32, _i64, _f64, _f64, _f64, _u64, _u32, _i64, _p]),
    "cymf_relmf_cells_host": (C.c_int, [_u64, _u32, _i64, _i64, _i32, _i32, _p]),
    "cymf_relmf_replay_epoch_dev": (C.c_int, [C.POINTER(Factors), C.c_int, _p, _i64, _p, _p, _p, _p,
                                              _i32, _i32, _i32, _i32, _f64, _f64, _f64, _p]),
    "cymf_relmf_fit_host": (C.c_int, [_p, _p, _i32, _i32, _i32, _p, _p, _p, _p, _i32, _f64, _f64, _f64,
                                      C.c_int, C.c_int, _u64]),
    "cymf_pack_rows_dev": (C.c_int, [_p, _p, C.c_int, _i64, _i32, _i32, _p]),
    "cymf_unpack_rows_dev": (C.c_int, [_p, _p, C.c_int, _i64, _i32, _i32, _p]),
    "cymf_fill_dev": (C.c_int, [_p, C.c_int, _i64, _f64, _p]),
    "cymf_bpr_hogwild_epoch_dev": (C.c_int, [C.POINTER(Factors), C.c_int, C.c_int, C.c_int, _p, _p, _i64, _p, _p,
                                             _i32, _i32, _i32, _i32, _f64, _f64, _u64, _u32, _i64, _p, _p]),
    "cymf_bpr_hogwild_range_dev": (C.c_int, [C.POINTER(Factors), C.c_int, C.c_int, C.c_int, _p, _p, _i64, _p, _p,
                                             _i32, _i32, _i32, _i32, _f64, _f64, _u64, _u32, _i64, _p, _i64, _p]),
    "cymf_bpr_negatives_host": (C.c_int, [_u64, _u32, _i64, _i64, _u32, _p]),
    "cymf_bpr_replay_epoch_dev":(C.c_int, [C.POINTER(Factors), C.c_int, _p, _p, _p, _i64, _p, _p,
                                            _i32, _i32, _i32, _i32, _f64, _f64, _p, _p]),
    "cymf_bpr_fit_host": (C.c_int, [_p, _p, _i32, _i32, _i32, _p, _p, _i64, _p, _p, _i32, _f64, _f64,
                                    C.c_int, C.c_int, _u64, _p]),
    "cymf_glove_hogwild_epoch_dev": (C.c_int, [C.POINTER(GloveParams), C.c_int, C.c_int, _p, _p, _p, _i64, _i32, _i32,
                                               _f64, _f64, _f64, _i64, _p, _p]),
    "cymf_glove_replay_epoch_dev": (C.c_int, [C.POINTER(GloveParams), _p, _p, _p, _i64, _i32, _i32, _f64, _f64, _f64,
                                              _p, _p]),
    "cymf_glove_fit_host": (C.c_int, [_p, _p, _p, _i64, _p, _p, _p, _p, _i64, _i64, _i32, _i32, _f64, _f64, _f64,
                                      C.c_int, _p]),
    "cymf_scan_workspace_bytes": (_i64, [_i64]),
    "cymf_exclusive_scan_u32_dev": (C.c_int, [_p, _p, _i64, _p, _p]),
    "cymf_sort_workspace_bytes": (_i64, [_i64]),
    "cymf_sort_pairs_dev": (C.c_int, [_p, _p, _i64, _i32, _p, _p]),
    "cymf_csr_transpose_workspace_bytes": (_i64, [_i64, _i64, _i64]),
    "cymf_csr_transpose_dev": (C.c_int, [_p, _p, _i64, _i64, _i64, _p, _p, _p, _p]),
    "cymf_deal_workspace_bytes": (_i64, [_i64]),
    "cymf_deal_rows_dev": (C.c_int, [_p, _i64, _i32, _i64, _p, _p, _p, _p]),
    "cymf_csr_block_workspace_bytes": (_i64, [_i64]),
    "cymf_csr_block_dev": (C.c_int, [_p, _p, _p, _i64, _p, _p, _p, _p, _p]),
    "cymf_word2vec_write_host": (C.c_int, [C.c_char_p, _p, _i64, _i32, _p]),
    "cymf_cooc_workspace_bytes": (_i64, [_i64, _i32]),
    "cymf_cooc_count_dev": (C.c_int, [_p, _p, _i64, _i32, _i32, _p, _p, _p, _i64, _p, _p, _p]),
    "cymf_gram_workspace_doubles": (_i64, [_i64, _i32]),
    "cymf_gram_dev": (C.c_int, [_p, C.c_int, _i64, _i32, _i32, _f64, C.c_int, _p, _i64, _p, _p, _p]),
    "cymf_gram_finalize_dev": (C.c_int, [_p, C.c_int, _i32, _i32, _f64, _p, _p]),
    "cymf_gram_sum_dev": (C.c_int, [C.POINTER(_p), _i32, _i32, _f64, _p, _p]),
    "cymf_als_cg_dev": (C.c_int, [_p, _p, _p, _i32, _p, _p, _p, _p, C.c_int, _i32, _i32, _f64, _f64, _i32, _i32, _i32,
                                  _p, _p, _p]),
    "cymf_als_rows_tc_dev": (C.c_int, [_p, _p, _p, _i32, _p, _p, C.c_int, _i32, _i32, _f64, _f64, _i32, _p, _p, _p]),
    "cymf_als_rows_dual_dev": (C.c_int, [_p, _p, _p, _i32, _i32, _i32, _p, _p, C.c_int, _i32, _i32, _f64, _f64, _i32, _p,
                                         _p, _p]),
    "cymf_als_ws_ctas": (_i32, []),
    "cymf_als_ws_schedule_host": (C.c_int, [_p, _p, _i32, _i32, _i32, _p, _p]),
    "cymf_als_rows_ws_dev": (C.c_int, [_p, _p, _i32, _p, _p, _p, _i64, C.c_int, _i32, _i32, _f64, _f64, _i32, _p, _p, _p]),
    "cymf_spd_inverse_dev": (C.c_int, [_p, _i32, _i32, _f64, C.c_int, _p, _p]),
    "cymf_chol_transforms_dev": (C.c_int, [_p, _i32, _i32, _f64, C.c_int, _p, _p, _p, _p, _p]),
    "cymf_rows_times_matrix_dev": (C.c_int, [_p, _p, _p, C.c_int, _i64, _i32, _p]),
    "cymf_rows_times_matrix_multi_dev": (C.c_int, [_p, C.POINTER(_p), _i32, _p, C.c_int, _i64, _i32, _p]),
    "cymf_als_heavy_workspace_doubles": (_i64, [_i64, _i32, _i32]),
    "cymf_als_heavy_rows_dev": (C.c_int, [_p, _p, _p, _i32, _p, _i32, _p, _p, _p, _f64, C.c_int, _i32, _i32, _f64, _p,
                                          _i64, _p]),
    "cymf_als_row_classes": (C.c_int, [_p, _i64, C.c_int, _i32, _p, _p]),
    "cymf_als_half_host": (C.c_int, [_p, _p, _p, _p, _i64, _i64, _i32, _f64, _f64, C.c_int, _f64, _i32, _p]),
    "cymf_eval_candidates_host": (C.c_int, [_i32, _i32, _p, _p, _p, _p, _i32, _u32, _p, _p, _i64]),
    "cymf_eval_rank_dev": (C.c_int, [_p, _p, _i32, _i32, _p, _p, _p, _i32, _p, _i32, _p, _i32, _p, _p, _p]),
}

_lib = None


class CymfError(RuntimeError):
    pass


def lib():
    """The loaded shared library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CymfError(f"{LIB_PATH} is missing: build the CUDA extension first "
                            "(python -c 'import __graft_entry__ as g; g.build()'); there is no CPU fallback")
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        if handle.cymf_abi_version() != 1:
            raise CymfError("libcymf_b200.so ABI version mismatch")
        _lib = handle
    return _lib


def check(status):
    if status != 0:
        raise CymfError(f"cymf_b200 status {status}: {lib().cymf_last_error().decode(errors='replace')}")


def require_cuda():
    import torch
    if not torch.cuda.is_available():
        raise CymfError("no CUDA device visible: cymf_b200 runs on B200 (sm_100a) only and has no CPU fallback")
    return torch


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def launch_count():
    return int(lib().cymf_launch_count())


def ld_for(K):
    """Device row stride: K rounded up to a multiple of 4 elements (16 B / 32 B vector slots)."""
    return (int(K) + 3) // 4 * 4


def ptr(t):
    """Device (or host) address of a torch tensor / numpy array, or NULL."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return C.c_void_p(t.data_ptr())
    return C.c_void_p(t.ctypes.data)


class HostRng:
    """The reference's UniformGenerator(0, n, seed) stream (cymf/math.pyx:12-18), generated on the host."""

    def __init__(self, seed=1234):
        self._h = lib().cymf_rng_create(seed)
        if not self._h:
            raise MemoryError("cymf_rng_create")

    def below(self, n, count):
        import numpy as np
        out = np.empty(int(count), np.int32)
        check(lib().cymf_rng_fill_below(self._h, int(n), out.ctypes.data_as(C.c_void_p), int(count)))
        return out

    def below64(self, n, count):
        import numpy as np
        out = np.empty(int(count), np.int64)
        check(lib().cymf_rng_fill_below64(self._h, int(n), out.ctypes.data_as(C.c_void_p), int(count)))
        return out

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.cymf_rng_destroy(self._h)
            self._h = None


def upload_factor(host_f64, dtype, device, ld=None):
    """Dense f64 [rows, K] ndarray -> device [rows, ld] tensor of `dtype` with zeroed pad columns."""
    import torch
    rows, K = host_f64.shape
    ld = ld_for(K) if ld is None else int(ld)
    src = torch.from_numpy(host_f64).to(device, non_blocking=True)
    dst = torch.empty((rows, ld), dtype=torch.float32 if dtype == F32 else torch.float64, device=device)
    check(lib().cymf_pack_rows_dev(ptr(src), ptr(dst), dtype, rows, K, ld, stream_ptr()))
    return dst


def download_factor(dev, K, out_f64):
    """Device [rows, ld] tensor -> the caller's f64 ndarray, in place (the reference mutates W/H in place)."""
    import torch
    rows, ld = dev.shape
    dtype = F32 if dev.dtype == torch.float32 else F64
    tmp = torch.empty((rows, K), dtype=torch.float64, device=dev.device)
    check(lib().cymf_unpack_rows_dev(ptr(dev), ptr(tmp), dtype, rows, K, ld, stream_ptr()))
    if out_f64.flags.c_contiguous and out_f64.dtype == "float64":
        torch.from_numpy(out_f64).copy_(tmp)          # straight D2H into the caller's buffer (DMA if it is pinned)
    else:
        out_f64[...] = tmp.cpu().numpy()
