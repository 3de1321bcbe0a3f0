"""ctypes binding of libsalg_b200.so (the C ABI declared in include/salg.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is present when a
context is created, this module raises.  Build the library with ``python -c "import __graft_entry__
as g; g.build()"`` (or ``make -C single-algebra_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# SALG_LIB_PATH: load an experiment build of the same library instead (csrc/Makefile: EXTRA / BUILD / OUT)
LIB_PATH = os.environ.get("SALG_LIB_PATH") and os.path.abspath(os.environ["SALG_LIB_PATH"]) or os.path.join(_HERE, "libsalg_b200.so")

OK, ERR_BAD_ARG, ERR_MASK_LEN, ERR_NOT_FITTED, ERR_CUDA, ERR_NCCL, ERR_NUMERIC, ERR_UNSUPPORTED, ERR_OOM = range(9)
F32, F64 = 0, 1
ROW, COLUMN = 0, 1
SVD_LANCZOS, SVD_RANDOM = 0, 1
NORM_QR, NORM_LU, NORM_NONE = 0, 1, 2
TRANSFORM_EXACT, TRANSFORM_REFERENCE_COMPAT = 0, 1


class SalgError(RuntimeError):
    """Non-zero status from the C ABI; ``code`` is the salg_status value, the message is the
    reference's own error string where it has one."""

    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


class PcaParams(C.Structure):
    _fields_ = [
        ("n_components", C.c_int32), ("svd_method", C.c_int32), ("n_oversamples", C.c_int32),
        ("n_power_iterations", C.c_int32), ("normalizer", C.c_int32), ("center", C.c_int32),
        ("verbose", C.c_int32), ("random_seed", C.c_uint32), ("alpha", C.c_double),
        ("tolerance", C.c_double), ("lanczos_max_steps", C.c_int32), ("keep_scores", C.c_int32),
        ("reserved", C.c_int32 * 6),
    ]


_P = C.c_void_p
_i64 = C.c_int64
_int = C.c_int

# name -> argtypes (all return int unless listed in _SPECIAL)
PROTOTYPES = {
    "salg_device_count": [C.POINTER(_int)],
    "salg_ctx_create": [_int, C.POINTER(_P)],
    "salg_nccl_unique_id": [_P],
    "salg_ctx_create_dist": [_int, _int, _int, _P, C.POINTER(_P)],
    "salg_ctx_destroy": [_P],
    "salg_ctx_sync": [_P],
    "salg_ctx_rank": [_P, C.POINTER(_int), C.POINTER(_int)],
    "salg_timer_start": [_P],
    "salg_timer_stop": [_P, C.POINTER(C.c_double)],
    "salg_ctx_set_spmm_impl": [_P, _int],
    "salg_launch_count": [_P, C.POINTER(_i64)],
    "salg_prof_enable": [_P, _int],
    "salg_prof_reset": [_P],
    "salg_prof_get": [_P, _int, C.POINTER(C.c_double), C.POINTER(_i64), C.POINTER(C.c_double)],
    "salg_csr_upload_f32": [_P, _i64, _i64, _i64, _P, _P, _P, C.POINTER(_P)],
    "salg_csr_upload_f64": [_P, _i64, _i64, _i64, _P, _P, _P, C.POINTER(_P)],
    "salg_csr_upload_i32_f32": [_P, _i64, _i64, _i64, _P, _P, _P, C.POINTER(_P)],
    "salg_csr_upload_i32_f64": [_P, _i64, _i64, _i64, _P, _P, _P, C.POINTER(_P)],
    "salg_csr_free": [_P],
    "salg_csr_dims": [_P, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_int)],
    "salg_csr_download_f32": [_P, _P, _P, _P, _P],
    "salg_csr_download_f64": [_P, _P, _P, _P, _P],
    "salg_csr_download_raw": [_P, _P, _P, _P, _P],
    "salg_csr_select_columns": [_P, _P, _P, _i64, C.POINTER(_P)],
    "salg_csr_transpose": [_P, _P, C.POINTER(_P)],
    "salg_csr_values_clone": [_P, _P, C.POINTER(_P)],
    "salg_csr_values_restore": [_P, _P, _P],
    "salg_dev_free": [_P, _P],
    "salg_nonzero_row": [_P, _P, _P],
    "salg_nonzero_col": [_P, _P, _P],
    "salg_csr_synth": [_P, _int, C.c_uint64, _i64, _i64, _i64, C.c_int32, _P, _P, _P, C.POINTER(_P)],
    "salg_sum_col_f32": [_P, _P, _P, _P],
    "salg_sum_col_f64": [_P, _P, _P, _P],
    "salg_sum_row_f32": [_P, _P, _P],
    "salg_sum_row_f64": [_P, _P, _P],
    "salg_col_stats_f64": [_P, _P, _P, _P, _P, _P],
    "salg_normalize_f32": [_P, _P, _P, _i64, C.c_float, _int],
    "salg_normalize_f64": [_P, _P, _P, _i64, C.c_double, _int],
    "salg_normalize_f32_u64": [_P, _P, _P, _i64, C.c_double, _int],
    "salg_log1p": [_P, _P],
    "salg_csc_upload_f32": [_P, _i64, _i64, _i64, _P, _P, _P, C.POINTER(_P)],
    "salg_csc_upload_f64": [_P, _i64, _i64, _i64, _P, _P, _P, C.POINTER(_P)],
    "salg_csc_sum_col_f32": [_P, _P, _P, _P],
    "salg_csc_sum_col_f64": [_P, _P, _P, _P],
    "salg_csc_sum_row_f32": [_P, _P, _P],
    "salg_csc_sum_row_f64": [_P, _P, _P],
    "salg_csc_normalize_f32": [_P, _P, _P, _i64, C.c_float, _int],
    "salg_csc_normalize_f64": [_P, _P, _P, _i64, C.c_double, _int],
    "salg_csc_normalize_f32_u64": [_P, _P, _P, _i64, C.c_double, _int],
    "salg_preprocess_f32": [_P, _P, C.c_float, _P, _P],
    "salg_preprocess_f64": [_P, _P, C.c_double, _P, _P],
    "salg_pca_params_default": [C.POINTER(PcaParams)],
    "salg_pca_fit_f32": [_P, _P, C.POINTER(PcaParams), _P, _i64, _P, _i64, _i64, C.POINTER(_P)],
    "salg_pca_fit_host_f32": [_P, _i64, _i64, _i64, _P, _P, _P, C.POINTER(PcaParams), _P, _i64, _P, _i64, _i64, C.POINTER(_P)],
    "salg_pca_fit_f64": [_P, _P, C.POINTER(PcaParams), _P, _i64, _P, _i64, _i64, C.POINTER(_P)],
    "salg_pca_free": [_P],
    "salg_pca_dims": [_P, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_int)],
    "salg_pca_components_f32": [_P, _P],
    "salg_pca_components_f64": [_P, _P],
    "salg_pca_singular_values_f64": [_P, _P],
    "salg_pca_explained_variance_f64": [_P, _P],
    "salg_pca_mean_f64": [_P, _P],
    "salg_pca_total_var": [_P, C.POINTER(C.c_double)],
    "salg_pca_n_samples": [_P, C.POINTER(_i64)],
    "salg_pca_numeric_flags": [_P, C.POINTER(_int)],
    "salg_pca_transform_f32": [_P, _P, _P, _int, _P],
    "salg_pca_transform_f64": [_P, _P, _P, _int, _P],
    "salg_pca_fit_scores_f32": [_P, _P, _P],
    "salg_pca_fit_scores_f64": [_P, _P, _P],
    "salg_pca_transform_device": [_P, _P, _P, _int],
    "salg_op_spmm_f32": [_P, _P, _int, _P, _i64, _P, _P],
    "salg_op_spmm_f64": [_P, _P, _int, _P, _i64, _P, _P],
    "salg_op_cholqr2_f32": [_P, _P, _i64, _i64, _P, _P],
    "salg_op_cholqr2_f64": [_P, _P, _i64, _i64, _P, _P],
    "salg_op_small_svd": [_P, _P, _i64, _P, _P, _P],
    "salg_op_spmm_bench": [_P, _P, _int, _i64, _int, C.POINTER(C.c_double)],
    "salg_op_tall_gram_f32": [_P, _P, _i64, _i64, _i64, _P, _P, _int, C.POINTER(C.c_double)],
}
_SPECIAL = {
    "salg_last_error": ([], C.c_char_p),
    "salg_version": ([], _int),
    "salg_prof_count": ([], _int),
    "salg_prof_name": ([_int], C.c_char_p),
}

_lib = None


def load():
    """Load (once) and return the ctypes library.  Raises if the extension has not been built —
    the product path has no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(run `python -c 'import __graft_entry__ as g; g.build()'`). "
            "single-algebra_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, argtypes in PROTOTYPES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _int
    for name, (argtypes, restype) in _SPECIAL.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def check(status):
    if status != OK:
        msg = load().salg_last_error()
        raise SalgError(status, msg.decode("utf-8", "replace") if msg else f"salg status {status}")


def ptr(a):
    """Raw data pointer of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
    return C.c_void_p(a.ctypes.data)
