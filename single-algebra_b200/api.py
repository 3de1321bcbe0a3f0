"""Host-side mirror of the reference's operator interface for the sparse-PCA hot path, over the C ABI.

Same names, argument meaning and error behaviour as the Rust surface (the Rust toolchain is absent in
this image, so the facade a Rust caller would use is shown in INTEGRATION.md and this module plays its
role for tests and benchmarks):

  single_algebra::dimred::pca::{SparsePCA, SparsePCABuilder, MaskedSparsePCA, MaskedSparsePCABuilder,
      SVDMethod, PowerIterationNormalizer}                (src/dimred/pca/mod.rs:37-62)
  single_algebra::sparse::MatrixSum::{sum_col, sum_col_squared, sum_row}  (src/sparse/mod.rs:67-102)
  single_algebra::{Normalize, Log1P}                       (src/utils/mod.rs:6-17)
  nalgebra_sparse::CsrMatrix<T>                            (the container all of them take)

Every method body is one or two calls into libsalg_b200.so; nothing here computes on the CPU.
Errors surface as `SalgError` carrying the reference's message (anyhow::Result in Rust).
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import os
from typing import Optional

import numpy as np

from . import _native as N
from ._native import SalgError


# --------------------------------------------------------------------------------------------------
# context
# --------------------------------------------------------------------------------------------------
class Context:
    """One GPU (`salg_ctx`): device, stream, workspaces and, for row-sharded runs, the NCCL
    communicator.  The reference's analogue is the ambient Rayon pool."""

    def __init__(self, device: int = 0, rank: int = 0, nranks: int = 1, nccl_unique_id: Optional[bytes] = None):
        lib = N.load()
        h = C.c_void_p()
        if nranks > 1:
            buf = C.create_string_buffer(nccl_unique_id, 128)
            N.check(lib.salg_ctx_create_dist(device, rank, nranks, buf, C.byref(h)))
        else:
            N.check(lib.salg_ctx_create(device, C.byref(h)))
        self._h = h
        self.device, self.rank, self.nranks = device, rank, nranks

    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        N.check(N.load().salg_nccl_unique_id(buf))
        return buf.raw

    def sync(self):
        N.check(N.load().salg_ctx_sync(self._h))

    def timer_start(self):
        N.check(N.load().salg_timer_start(self._h))

    def timer_stop(self) -> float:
        ms = C.c_double()
        N.check(N.load().salg_timer_stop(self._h, C.byref(ms)))
        return ms.value

    def set_spmm_impl(self, impl: str):
        """'tm' (tcgen05, sparse operand expanded into TMEM: default for f32), 'tc' (tcgen05, dense tile in shared memory)
        or 'chunk' (CUDA-core kernels)."""
        N.check(N.load().salg_ctx_set_spmm_impl(self._h, {'tc': 0, 'chunk': 1, 'tm': 2}[impl]))

    def launch_count(self) -> int:
        n = C.c_int64()
        N.check(N.load().salg_launch_count(self._h, C.byref(n)))
        return n.value

    def prof_enable(self, on=True, products_only=False):
        """Per-class CUDA-event timing; `products_only` times just the two sparse-product classes (two event records
        per product instead of two per scope: the small-side chain of a fit is launch-bound)."""
        N.check(N.load().salg_prof_enable(self._h, (2 if products_only else 1) if on else 0))

    def prof_reset(self):
        N.check(N.load().salg_prof_reset(self._h))

    def prof(self):
        """{class name: (device ms, launches, algorithmic bytes)} since the last reset."""
        lib = N.load()
        out = {}
        for i in range(lib.salg_prof_count()):
            ms, n, b = C.c_double(), C.c_int64(), C.c_double()
            N.check(lib.salg_prof_get(self._h, i, C.byref(ms), C.byref(n), C.byref(b)))
            if n.value:
                out[lib.salg_prof_name(i).decode()] = (ms.value, n.value, b.value)
        return out

    def close(self):
        if self._h:
            N.load().salg_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    """Process-wide context on cuda:LOCAL_RANK (created on first use)."""
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(int(os.environ.get("LOCAL_RANK", "0")))
    return _default_ctx


def set_default_context(ctx: Optional[Context]):
    global _default_ctx
    _default_ctx = ctx


def device_count() -> int:
    n = C.c_int()
    N.check(N.load().salg_device_count(C.byref(n)))
    return n.value


# --------------------------------------------------------------------------------------------------
# enums of the reference
# --------------------------------------------------------------------------------------------------
class Direction:
    """single_utilities::types::Direction"""
    ROW = N.ROW
    COLUMN = N.COLUMN


class PowerIterationNormalizer:
    """single_svdlib::randomized::PowerIterationNormalizer (re-exported at src/dimred/pca/mod.rs:41)."""
    QR = N.NORM_QR
    LU = N.NORM_LU
    NoNormalization = N.NORM_NONE


@dataclasses.dataclass(frozen=True)
class SVDMethod:
    """src/dimred/pca/mod.rs:49-62.  `SVDMethod.Lanczos` (the default, :64-68) or
    `SVDMethod.Random(n_oversamples, n_power_iterations, normalizer)`."""
    kind: int = N.SVD_LANCZOS
    n_oversamples: int = 10
    n_power_iterations: int = 7
    normalizer: int = N.NORM_QR

    @staticmethod
    def Random(n_oversamples: int, n_power_iterations: int, normalizer: int = N.NORM_QR) -> "SVDMethod":
        return SVDMethod(N.SVD_RANDOM, n_oversamples, n_power_iterations, normalizer)

    @staticmethod
    def default() -> "SVDMethod":
        return SVDMethod()


SVDMethod.Lanczos = SVDMethod()


# --------------------------------------------------------------------------------------------------
# CSR containers
# --------------------------------------------------------------------------------------------------
class DeviceCsr:
    """Device-resident CSR row shard (`salg_csr`).  Created by `CsrMatrix.to_device`, `synth` or
    `select_columns`; value type fixed at creation."""

    def __init__(self, handle, ctx: Context):
        self._h = handle
        self.ctx = ctx
        r, c, z, d = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int()
        N.check(N.load().salg_csr_dims(handle, C.byref(r), C.byref(c), C.byref(z), C.byref(d)))
        self.nrows, self.ncols, self.nnz = r.value, c.value, z.value
        self.dtype = np.dtype(np.float64 if d.value == N.F64 else np.float32)

    @property
    def _sfx(self):
        return "f64" if self.dtype == np.float64 else "f32"

    def download(self):
        """(row_offsets u64, col_indices u64, values T) — nalgebra-sparse layout."""
        off = np.empty(self.nrows + 1, np.uint64)
        idx = np.empty(self.nnz, np.uint64)
        val = np.empty(self.nnz, self.dtype)
        fn = getattr(N.load(), f"salg_csr_download_{self._sfx}")
        N.check(fn(self.ctx._h, self._h, N.ptr(off), N.ptr(idx), N.ptr(val)))
        return off, idx, val

    def download_raw(self, off=None, idx=None, val=None):
        """Device layout into caller buffers (e.g. pinned): int64 offsets, uint32 indices, values."""
        N.check(N.load().salg_csr_download_raw(self.ctx._h, self._h, N.ptr(off), N.ptr(idx), N.ptr(val)))

    def download_values(self):
        val = np.empty(self.nnz, self.dtype)
        fn = getattr(N.load(), f"salg_csr_download_{self._sfx}")
        N.check(fn(self.ctx._h, self._h, None, None, N.ptr(val)))
        return val

    def select_columns(self, mask) -> "DeviceCsr":
        """MaskedCSRMatrix::new(x, mask) as a materialised compaction (pca/sparse_masked/mod.rs:313)."""
        m = np.ascontiguousarray(np.asarray(mask, dtype=bool).astype(np.uint8))
        h = C.c_void_p()
        N.check(N.load().salg_csr_select_columns(self.ctx._h, self._h, N.ptr(m), len(m), C.byref(h)))
        return DeviceCsr(h, self.ctx)

    def transpose(self) -> "DeviceCsr":
        """New handle holding the CSR of A^T (= the CSC arrays of A); device-side stable sort by column."""
        h = C.c_void_p()
        N.check(N.load().salg_csr_transpose(self.ctx._h, self._h, C.byref(h)))
        return DeviceCsr(h, self.ctx)

    def clone_values(self):
        """Opaque device copy of the value array (restore_values puts it back: the in-place Normalize / Log1P chain can
        then be repeated from the same raw counts without another upload)."""
        h = C.c_void_p()
        N.check(N.load().salg_csr_values_clone(self.ctx._h, self._h, C.byref(h)))
        return h

    def restore_values(self, clone):
        N.check(N.load().salg_csr_values_restore(self.ctx._h, self._h, clone))

    def free_values_clone(self, clone):
        N.check(N.load().salg_dev_free(self.ctx._h, clone))

    # MatrixNonZero (src/sparse/csr.rs:23-122) -------------------------------------------------------
    def nonzero_row(self):
        out = np.empty(self.nrows, np.uint64)
        N.check(N.load().salg_nonzero_row(self.ctx._h, self._h, N.ptr(out)))
        return out

    def nonzero_col(self):
        out = np.empty(self.ncols, np.uint64)
        N.check(N.load().salg_nonzero_col(self.ctx._h, self._h, N.ptr(out)))
        return out

    # MatrixSum ------------------------------------------------------------------------------------
    def sum_col(self):
        out = np.empty(self.ncols, self.dtype)
        N.check(getattr(N.load(), f"salg_sum_col_{self._sfx}")(self.ctx._h, self._h, N.ptr(out), None))
        return out

    def sum_col_and_squared(self):
        s = np.empty(self.ncols, self.dtype)
        q = np.empty(self.ncols, self.dtype)
        N.check(getattr(N.load(), f"salg_sum_col_{self._sfx}")(self.ctx._h, self._h, N.ptr(s), N.ptr(q)))
        return s, q

    def sum_col_squared(self):
        return self.sum_col_and_squared()[1]

    def sum_row(self):
        out = np.empty(self.nrows, self.dtype)
        N.check(getattr(N.load(), f"salg_sum_row_{self._sfx}")(self.ctx._h, self._h, N.ptr(out)))
        return out

    def col_stats(self):
        """(sum, sumsq, nonzero_col, var_col) in f64 — SURVEY §8f-1."""
        a = [np.empty(self.ncols, np.float64) for _ in range(4)]
        N.check(N.load().salg_col_stats_f64(self.ctx._h, self._h, *[N.ptr(x) for x in a]))
        return tuple(a)

    # Normalize / Log1P ------------------------------------------------------------------------------
    def normalize(self, sums, target, direction):
        sums = np.ascontiguousarray(sums)
        lib = N.load()
        if self.dtype == np.float64:
            sums = sums.astype(np.float64, copy=False)
            N.check(lib.salg_normalize_f64(self.ctx._h, self._h, N.ptr(sums), len(sums), float(target), direction))
        elif sums.dtype == np.float64:
            N.check(lib.salg_normalize_f32_u64(self.ctx._h, self._h, N.ptr(sums), len(sums), float(target), direction))
        else:
            sums = sums.astype(np.float32, copy=False)
            N.check(lib.salg_normalize_f32(self.ctx._h, self._h, N.ptr(sums), len(sums), float(target), direction))

    def log1p_normalize(self):
        N.check(N.load().salg_log1p(self.ctx._h, self._h))

    def preprocess(self, target):
        """Fused sum_row -> normalize(ROW, target) -> log1p -> (sum_col, sum_col_squared)."""
        s = np.empty(self.ncols, self.dtype)
        q = np.empty(self.ncols, self.dtype)
        fn = getattr(N.load(), f"salg_preprocess_{self._sfx}")
        N.check(fn(self.ctx._h, self._h, float(target), N.ptr(s), N.ptr(q)))
        return s, q

    def preprocess_device(self, target):
        """The same fused chain with nothing copied back (the statistics are recomputed by the fit that follows)."""
        fn = getattr(N.load(), f"salg_preprocess_{self._sfx}")
        N.check(fn(self.ctx._h, self._h, float(target), None, None))

    def free(self):
        if self._h:
            N.load().salg_csr_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class CsrMatrix:
    """nalgebra_sparse::CsrMatrix<T> as the reference's functions receive it: host arrays
    `row_offsets` / `col_indices` (usize) and `values` (T).  The MatrixSum / Normalize / Log1P trait
    methods are implemented on it exactly as in src/sparse/csr.rs, each body being an FFI call; the
    in-place traits refresh the host `values` so `&mut self` semantics hold."""

    def __init__(self, nrows, ncols, row_offsets, col_indices, values, ctx: Optional[Context] = None):
        self.nrows, self.ncols = int(nrows), int(ncols)
        self.row_offsets = np.ascontiguousarray(row_offsets, dtype=np.uint64)
        ci = np.asarray(col_indices)
        # u64 is nalgebra's layout; int32 (scipy/AnnData) is accepted without the usize detour
        self.col_indices = np.ascontiguousarray(ci if ci.dtype == np.int32 else ci.astype(np.uint64, copy=False))
        v = np.asarray(values)
        if v.dtype not in (np.float32, np.float64):
            v = v.astype(np.float64)
        self.values = np.ascontiguousarray(v)
        self._ctx = ctx
        self._dev: Optional[DeviceCsr] = None

    @classmethod
    def from_scipy(cls, A, ctx=None):
        A = A.tocsr()
        A.sort_indices()
        return cls(A.shape[0], A.shape[1], A.indptr, A.indices.astype(np.uint64), A.data, ctx)

    @property
    def dtype(self):
        return self.values.dtype

    @property
    def nnz(self):
        return len(self.values)

    @property
    def ctx(self):
        return self._ctx or default_context()

    def to_device(self, ctx: Optional[Context] = None) -> DeviceCsr:
        """Upload (usize -> u32 narrowing + validation on the device); cached until values change."""
        ctx = ctx or self.ctx
        if self._dev is not None and self._dev.ctx is ctx and self._dev._h:
            return self._dev
        lib = N.load()
        h = C.c_void_p()
        sfx = "f64" if self.dtype == np.float64 else "f32"
        if self.col_indices.dtype == np.int32:
            off = self.row_offsets.view(np.int64)
            fn = getattr(lib, f"salg_csr_upload_i32_{sfx}")
        else:
            off = self.row_offsets
            fn = getattr(lib, f"salg_csr_upload_{sfx}")
        N.check(fn(ctx._h, self.nrows, self.ncols, self.nnz, N.ptr(off), N.ptr(self.col_indices),
                   N.ptr(self.values), C.byref(h)))
        self._dev = DeviceCsr(h, ctx)
        return self._dev

    def drop_device(self):
        if self._dev is not None:
            self._dev.free()
            self._dev = None

    # MatrixSum (src/sparse/csr.rs:259-312, 314-392, 558-608) ------------------------------------------
    def sum_col(self):
        return self.to_device().sum_col()

    def sum_col_squared(self):
        return self.to_device().sum_col_squared()

    def sum_row(self):
        return self.to_device().sum_row()

    # MatrixNonZero (src/sparse/csr.rs:23-122) ----------------------------------------------------------
    def nonzero_col(self):
        return self.to_device().nonzero_col()

    def nonzero_row(self):
        return self.to_device().nonzero_row()

    # Normalize / Log1P (src/sparse/csr.rs:1013-1079) ------------------------------------------------------
    def normalize(self, sums, target, direction):
        d = self.to_device()
        d.normalize(sums, target, direction)
        self.values = d.download_values()

    def log1p_normalize(self):
        d = self.to_device()
        d.log1p_normalize()
        self.values = d.download_values()


class CscMatrix:
    """nalgebra_sparse::CscMatrix<T> (host `col_offsets` / `row_indices` (usize) / `values`) with the reference's
    MatrixSum / Normalize / Log1P implementations for it (src/sparse/csc.rs:157-220, 323-335, 680-746).  On the device
    it is the CSR of A^T, so a column of A is a stored row; every method body is one FFI call."""

    def __init__(self, nrows, ncols, col_offsets, row_indices, values, ctx: Optional[Context] = None):
        self.nrows, self.ncols = int(nrows), int(ncols)
        self.col_offsets = np.ascontiguousarray(col_offsets, dtype=np.uint64)
        self.row_indices = np.ascontiguousarray(np.asarray(row_indices).astype(np.uint64, copy=False))
        v = np.asarray(values)
        if v.dtype not in (np.float32, np.float64):
            v = v.astype(np.float64)
        self.values = np.ascontiguousarray(v)
        self._ctx = ctx
        self._h = None

    @classmethod
    def from_scipy(cls, A, ctx=None):
        A = A.tocsc()
        A.sort_indices()
        return cls(A.shape[0], A.shape[1], A.indptr, A.indices.astype(np.uint64), A.data, ctx)

    @property
    def dtype(self):
        return self.values.dtype

    @property
    def nnz(self):
        return len(self.values)

    @property
    def ctx(self):
        return self._ctx or default_context()

    @property
    def _sfx(self):
        return "f64" if self.dtype == np.float64 else "f32"

    def to_device(self):
        if self._h is None:
            h = C.c_void_p()
            N.check(getattr(N.load(), f"salg_csc_upload_{self._sfx}")(
                self.ctx._h, self.nrows, self.ncols, self.nnz, N.ptr(self.col_offsets), N.ptr(self.row_indices),
                N.ptr(self.values), C.byref(h)))
            self._h = h
        return self._h

    def drop_device(self):
        if self._h is not None:
            N.load().salg_csr_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.drop_device()
        except Exception:
            pass

    def _refresh_values(self):
        fn = getattr(N.load(), f"salg_csr_download_{self._sfx}")
        val = np.empty(self.nnz, self.dtype)
        N.check(fn(self.ctx._h, self._h, None, None, N.ptr(val)))
        self.values = val

    # MatrixSum (src/sparse/csc.rs:157-220, 323-335)
    def sum_col(self):
        out = np.empty(self.ncols, self.dtype)
        N.check(getattr(N.load(), f"salg_csc_sum_col_{self._sfx}")(self.ctx._h, self.to_device(), N.ptr(out), None))
        return out

    def sum_col_squared(self):
        out = np.empty(self.ncols, self.dtype)
        N.check(getattr(N.load(), f"salg_csc_sum_col_{self._sfx}")(self.ctx._h, self.to_device(), None, N.ptr(out)))
        return out

    def sum_row(self):
        out = np.empty(self.nrows, self.dtype)
        N.check(getattr(N.load(), f"salg_csc_sum_row_{self._sfx}")(self.ctx._h, self.to_device(), N.ptr(out)))
        return out

    # Normalize / Log1P (src/sparse/csc.rs:680-746)
    def normalize(self, sums, target, direction):
        sums = np.ascontiguousarray(sums)
        lib, h = N.load(), self.to_device()
        if self.dtype == np.float64:
            sums = sums.astype(np.float64, copy=False)
            N.check(lib.salg_csc_normalize_f64(self.ctx._h, h, N.ptr(sums), len(sums), float(target), direction))
        elif sums.dtype == np.float64:
            N.check(lib.salg_csc_normalize_f32_u64(self.ctx._h, h, N.ptr(sums), len(sums), float(target), direction))
        else:
            sums = sums.astype(np.float32, copy=False)
            N.check(lib.salg_csc_normalize_f32(self.ctx._h, h, N.ptr(sums), len(sums), float(target), direction))
        self._refresh_values()

    def log1p_normalize(self):
        N.check(N.load().salg_log1p(self.ctx._h, self.to_device()))
        self._refresh_values()


def _as_device(x, ctx=None) -> DeviceCsr:
    return x if isinstance(x, DeviceCsr) else x.to_device(ctx)


# --------------------------------------------------------------------------------------------------
# PCA
# --------------------------------------------------------------------------------------------------
class _PCABase:
    _masked = False

    def __init__(self, n_components, alpha, tolerance, random_seed, center, verbose, svdmethod, mask=None):
        self.n_components = int(n_components)
        self.alpha = alpha
        self.tolerance = 1e-6 if tolerance is None else tolerance
        self.random_seed = 42 if random_seed is None else int(random_seed)
        self.center = bool(center)
        self.verbose = bool(verbose)
        self.svdmethod = svdmethod
        self.mask = None if mask is None else np.asarray(mask, dtype=bool)
        self.components_ = None            # d x n_eff  (pca/sparse/mod.rs:208)
        self.explained_variance_ = None    # d          (:210-216)
        self.mean_ = None                  # ncols      (:106-117, masked :280-291)
        self.singular_values_ = None
        self.total_var_ = None
        self._model = None
        self._dtype = None
        # extensions outside the reference API (SURVEY §5 "config"): injected test matrix, transform
        # semantics, Lanczos step cap
        self.transform_mode = N.TRANSFORM_EXACT
        self.lanczos_max_steps = 0

    def _params(self, keep_scores):
        p = N.PcaParams()
        N.check(N.load().salg_pca_params_default(C.byref(p)))
        p.n_components = self.n_components
        p.svd_method = self.svdmethod.kind
        p.n_oversamples = self.svdmethod.n_oversamples
        p.n_power_iterations = self.svdmethod.n_power_iterations
        p.normalizer = self.svdmethod.normalizer
        p.center = int(self.center)
        p.verbose = int(self.verbose)
        p.random_seed = self.random_seed
        p.alpha = float(self.alpha)
        p.tolerance = float(self.tolerance)
        p.lanczos_max_steps = int(self.lanczos_max_steps)
        p.keep_scores = int(keep_scores)
        return p

    def _free_model(self):
        if self._model is not None:
            N.load().salg_pca_free(self._model)
            self._model = None

    def __del__(self):
        try:
            self._free_model()
        except Exception:
            pass

    def _fit_host(self, x, omega, keep_scores, fetch=True):
        """Host-resident `CsrMatrix` (int32 indices, f32): `salg_pca_fit_host_f32` — a masked randomized fit streams the
        matrix through the statistics + compaction pass while it uploads."""
        lib = N.load()
        ctx = x.ctx
        mask_arr, mask_len = None, 0
        if self._masked:
            mask_arr = np.ascontiguousarray(self.mask.astype(np.uint8))
            mask_len = len(mask_arr)
        om, orows, ocols = None, 0, 0
        if omega is not None:
            om = np.ascontiguousarray(omega, dtype=np.float32)
            orows, ocols = om.shape
        p = self._params(keep_scores)
        h = C.c_void_p()
        self._free_model()
        N.check(lib.salg_pca_fit_host_f32(ctx._h, x.nrows, x.ncols, x.nnz, N.ptr(x.row_offsets.view(np.int64)),
                                          N.ptr(x.col_indices), N.ptr(x.values), C.byref(p), N.ptr(mask_arr), mask_len,
                                          N.ptr(om), orows, ocols, C.byref(h)))
        self._model = h
        self._dtype = np.dtype(np.float32)
        self._ctx = ctx
        if fetch:
            self._fetch()
        return self

    def _fit(self, x, omega, keep_scores, fetch=True):
        lib = N.load()
        if (isinstance(x, CsrMatrix) and x._dev is None and x.dtype == np.float32 and x.col_indices.dtype == np.int32
                and self._masked and self.svdmethod.kind == N.SVD_RANDOM):
            return self._fit_host(x, omega, keep_scores, fetch)
        d = _as_device(x)
        sfx = "f64" if d.dtype == np.float64 else "f32"
        mask_arr, mask_len = None, 0
        if self._masked:
            mask_arr = np.ascontiguousarray(self.mask.astype(np.uint8))
            mask_len = len(mask_arr)
        om, orows, ocols = None, 0, 0
        if omega is not None:
            om = np.ascontiguousarray(omega, dtype=d.dtype)
            orows, ocols = om.shape
        p = self._params(keep_scores)
        h = C.c_void_p()
        self._free_model()
        N.check(getattr(lib, f"salg_pca_fit_{sfx}")(d.ctx._h, d._h, C.byref(p), N.ptr(mask_arr), mask_len,
                                                    N.ptr(om), orows, ocols, C.byref(h)))
        self._model = h
        self._dtype = d.dtype
        self._ctx = d.ctx
        flags = self.numeric_flags()
        if flags & 4:
            import warnings
            warnings.warn("Lanczos stopped at its step limit before every requested singular triplet met the acceptance "
                          "bound; the returned components are the best available (numeric_flags() & 4)", RuntimeWarning,
                          stacklevel=3)
        if fetch:
            self._fetch()
        return self

    def _fetch(self):
        lib = N.load()
        h = self._model
        dd, ne, nc, dt = C.c_int64(), C.c_int64(), C.c_int64(), C.c_int()
        N.check(lib.salg_pca_dims(h, C.byref(dd), C.byref(ne), C.byref(nc), C.byref(dt)))
        d, n_eff, ncols = dd.value, ne.value, nc.value
        sfx = "f64" if self._dtype == np.float64 else "f32"
        comp = np.empty((d, n_eff), self._dtype)
        N.check(getattr(lib, f"salg_pca_components_{sfx}")(h, N.ptr(comp)))
        s = np.empty(d, np.float64)
        ev = np.empty(d, np.float64)
        mean = np.empty(ncols, np.float64)
        tv = C.c_double()
        N.check(lib.salg_pca_singular_values_f64(h, N.ptr(s)))
        N.check(lib.salg_pca_explained_variance_f64(h, N.ptr(ev)))
        N.check(lib.salg_pca_mean_f64(h, N.ptr(mean)))
        N.check(lib.salg_pca_total_var(h, C.byref(tv)))
        self.components_ = comp
        self.singular_values_ = s
        self.explained_variance_ = ev.astype(self._dtype)
        self.mean_ = mean.astype(self._dtype)
        self.total_var_ = tv.value
        ns = C.c_int64()
        N.check(lib.salg_pca_n_samples(h, C.byref(ns)))
        self._n_samples = ns.value

    # -- reference API ------------------------------------------------------------------------------------
    def fit(self, x, omega=None):
        """`fit(&mut self, x: &CsrMatrix<T>)` (pca/sparse/mod.rs:102, pca/sparse_masked/mod.rs:255).
        `omega` (extension): host Gaussian test matrix shared with the oracle for parity."""
        return self._fit(x, omega, keep_scores=False)

    def transform(self, x):
        """`transform(&self, x)` (pca/sparse/mod.rs:255, pca/sparse_masked/mod.rs:438)."""
        if self._masked and x.ncols != len(self.mask):      # checked before the fitted state (pca/sparse_masked/mod.rs:440-444)
            raise SalgError(N.ERR_MASK_LEN,
                            "The mask vector length and the number of features (columns) have to be the same!")
        if self._model is None:
            raise SalgError(N.ERR_NOT_FITTED, "Must be fitted before transform!")
        lib = N.load()
        d = _as_device(x)
        sfx = "f64" if d.dtype == np.float64 else "f32"
        out = np.empty((d.nrows, self.components_.shape[0]), d.dtype)
        N.check(getattr(lib, f"salg_pca_transform_{sfx}")(d.ctx._h, self._model, d._h, self.transform_mode, N.ptr(out)))
        return out

    def fit_transform(self, x, omega=None, out=None):
        """`fit_transform(&mut self, x)` (pca/sparse/mod.rs:355-358): fit, then the projection of the
        same rows; the projection runs inside the fit call while the compacted operator is resident."""
        if self.transform_mode != N.TRANSFORM_EXACT:
            # fit + transform literally (pca/sparse/mod.rs:355-358): the projection kept by the fit is the EXACT one
            self._fit(x, omega, keep_scores=False)
            sc = self.transform(x)
            if out is not None:
                out[...] = sc
                return out
            return sc
        self._fit(x, omega, keep_scores=True)
        lib = N.load()
        sfx = "f64" if self._dtype == np.float64 else "f32"
        nrows = x.nrows
        if out is None:
            out = np.empty((nrows, self.components_.shape[0]), self._dtype)
        assert out.shape == (nrows, self.components_.shape[0]) and out.dtype == self._dtype
        N.check(getattr(lib, f"salg_pca_fit_scores_{sfx}")(self._ctx._h, self._model, N.ptr(out)))   # `out` may be pinned
        return out

    def feature_importances(self):
        """pca/sparse/mod.rs:295-302 — squared loadings (host; negligible); unfitted: "Model must be fitted first!" (:299)."""
        if self.components_ is None:
            raise SalgError(N.ERR_NOT_FITTED, "Model must be fitted first!")
        return self.components_ * self.components_

    def explained_variance_ratio(self):
        """pca/sparse/mod.rs:312-322 — normalised by the sum over the COMPUTED components."""
        if self.explained_variance_ is None:
            raise SalgError(N.ERR_NOT_FITTED, "Model must be fitted first!")
        ev = self.explained_variance_
        return ev / ev.sum()

    def cumulative_explained_variance_ratio(self):
        """pca/sparse/mod.rs:333-343."""
        return np.cumsum(self.explained_variance_ratio())

    # what the reference only prints (pca/sparse/mod.rs:225-238) or normalises by the computed part
    def explained_variance_ratio_total(self):
        """explained_variance_ / total variance of the (kept, centred) columns — the scikit-learn definition; the
        reference divides by the sum over the computed components only (pca/sparse/mod.rs:318-319)."""
        if self.explained_variance_ is None:
            raise SalgError(N.ERR_NOT_FITTED, "Model must be fitted first!")
        return self.explained_variance_ / self.total_var_

    def noise_variance(self):
        """(total_var - sum explained) / (min(n_samples, n_features) - n_components): the value the reference prints under
        `verbose` (pca/sparse/mod.rs:225-238), returned; None when every component was computed."""
        if self.explained_variance_ is None:
            raise SalgError(N.ERR_NOT_FITTED, "Model must be fitted first!")
        min_dim = min(self._n_samples, self.components_.shape[1])
        d = len(self.explained_variance_)
        if d >= min_dim:
            return None
        return float((self.total_var_ - self.explained_variance_.sum()) / (min_dim - d))

    def numeric_flags(self):
        f = C.c_int()
        N.check(N.load().salg_pca_numeric_flags(self._model, C.byref(f)))
        return f.value


class SparsePCA(_PCABase):
    """src/dimred/pca/sparse/mod.rs:33-47; constructor :63-84."""

    def __init__(self, n_components, alpha, tollerance=None, random_seed=None, center=True, verbose=False,
                 svdmethod=SVDMethod()):
        super().__init__(n_components, alpha, tollerance, random_seed, center, verbose, svdmethod)


class MaskedSparsePCA(_PCABase):
    """src/dimred/pca/sparse_masked/mod.rs:179-194; constructor :214-237."""
    _masked = True

    def __init__(self, n_components, alpha, tolerance=None, random_seed=None, center=True, verbose=False,
                 mask=(), svdmethod=SVDMethod()):
        super().__init__(n_components, alpha, tolerance, random_seed, center, verbose, svdmethod, mask=mask)


class _BuilderBase:
    def __init__(self):
        # defaults: pca/sparse/mod.rs:388-403, pca/sparse_masked/mod.rs:51-67
        self._n_components = 50
        self._alpha = 1.0
        self._tolerance = 1e-6
        self._random_seed = 42
        self._center = True
        self._verbose = False
        self._svdmethod = SVDMethod.default()

    @classmethod
    def new(cls):
        return cls()

    def n_components(self, n):
        self._n_components = n
        return self

    def alpha(self, a):
        self._alpha = a
        return self

    def tolerance(self, t):
        self._tolerance = t
        return self

    def random_seed(self, s):
        self._random_seed = s
        return self

    def center(self, c):
        self._center = c
        return self

    def verbose(self, v):
        self._verbose = v
        return self

    def svd_method(self, m):
        self._svdmethod = m
        return self


class SparsePCABuilder(_BuilderBase):
    """src/dimred/pca/sparse/mod.rs:375-484."""

    def build(self) -> SparsePCA:
        return SparsePCA(self._n_components, self._alpha, self._tolerance, self._random_seed, self._center,
                         self._verbose, self._svdmethod)


class MaskedSparsePCABuilder(_BuilderBase):
    """src/dimred/pca/sparse_masked/mod.rs:37-160."""

    def __init__(self):
        super().__init__()
        self._mask = []

    def mask(self, mask):
        self._mask = mask
        return self

    def build(self) -> MaskedSparsePCA:
        return MaskedSparsePCA(self._n_components, self._alpha, self._tolerance, self._random_seed, self._center,
                               self._verbose, self._mask, self._svdmethod)


# --------------------------------------------------------------------------------------------------
# operator-level helpers (parity tests / microbenchmarks)
# --------------------------------------------------------------------------------------------------
def op_spmm(x, dense, mu=None, transposed=False):
    d = _as_device(x)
    dense = np.ascontiguousarray(dense, dtype=d.dtype)
    k = dense.shape[1]
    n_out = d.ncols if transposed else d.nrows
    out = np.empty((n_out, k), d.dtype)
    mu_a = None if mu is None else np.ascontiguousarray(mu, dtype=d.dtype)
    sfx = "f64" if d.dtype == np.float64 else "f32"
    N.check(getattr(N.load(), f"salg_op_spmm_{sfx}")(d.ctx._h, d._h, int(transposed), N.ptr(dense), k, N.ptr(mu_a),
                                                     N.ptr(out)))
    return out


def op_cholqr2(panel, ctx=None):
    ctx = ctx or default_context()
    panel = np.ascontiguousarray(panel)
    m, k = panel.shape
    q = np.empty_like(panel)
    r = np.empty((k, k), np.float64)
    sfx = "f64" if panel.dtype == np.float64 else "f32"
    N.check(getattr(N.load(), f"salg_op_cholqr2_{sfx}")(ctx._h, N.ptr(panel), m, k, N.ptr(q), N.ptr(r)))
    return q, r


def op_small_svd(a, ctx=None):
    ctx = ctx or default_context()
    a = np.ascontiguousarray(a, dtype=np.float64)
    k = a.shape[0]
    u, s, vt = np.empty((k, k)), np.empty(k), np.empty((k, k))
    N.check(N.load().salg_op_small_svd(ctx._h, N.ptr(a), k, N.ptr(u), N.ptr(s), N.ptr(vt)))
    return u, s, vt


def op_tall_gram(panel=None, ctx=None, device_rows=0, k=60, iters=0):
    """Gram matrix and column sums of a tall f32 panel through the fused tcgen05 pass (test / probe hook).
    Returns (gram k x k f64, colsum k f64, avg kernel ms)."""
    ctx = ctx or default_context()
    if panel is not None:
        panel = np.ascontiguousarray(panel, dtype=np.float32)
        m, k = panel.shape
    else:
        m = 0
    g = np.empty((k, k), np.float64)
    cs = np.empty(k, np.float64)
    ms = C.c_double()
    N.check(N.load().salg_op_tall_gram_f32(ctx._h, N.ptr(panel) if panel is not None else None, m, k, device_rows,
                                           N.ptr(g), N.ptr(cs), iters, C.byref(ms)))
    return g, cs, ms.value


def op_spmm_bench(x, transposed=False, k=60, iters=10):
    d = _as_device(x)
    ms = C.c_double()
    N.check(N.load().salg_op_spmm_bench(d.ctx._h, d._h, int(transposed), k, iters, C.byref(ms)))
    return ms.value


def synth_device(spec, row0=0, nrows=None, dtype=np.float32, ctx=None) -> DeviceCsr:
    """Rows [row0, row0+nrows) of the synthetic count matrix described by `spec` (synth.SynthSpec),
    generated on the device; bit-identical to synth.generate_rows on the host."""
    ctx = ctx or default_context()
    nrows = spec.nrows - row0 if nrows is None else nrows
    bl = np.ascontiguousarray(spec.base_level, dtype=np.uint8)
    sf = np.ascontiguousarray(spec.sf_offset, dtype=np.int32)
    cdf = np.ascontiguousarray(spec.cdf, dtype=np.uint32)
    h = C.c_void_p()
    N.check(N.load().salg_csr_synth(ctx._h, N.F64 if np.dtype(dtype) == np.float64 else N.F32, spec.seed, row0, nrows,
                                    spec.ncols, spec.n_clusters, N.ptr(bl), N.ptr(sf), N.ptr(cdf), C.byref(h)))
    return DeviceCsr(h, ctx)
