"""ctypes binding of oracle/_build/libcpuref.so (oracle/cpu_ref.cpp): the multi-threaded C++ restatement of the
reference's sparse-PCA path and the host-side synthetic generator.

TEST INFRASTRUCTURE ONLY — see the header of cpu_ref.cpp.  Imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libcpuref.so")
_lib = None


def build():
    subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.cpuref_max_threads.restype = C.c_int
        L.cpuref_set_threads.argtypes = [C.c_int]
        L.cpuref_synth.restype = C.c_int64
        L.cpuref_synth.argtypes = [C.c_uint64, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        for name in ("cpuref_pca_fit_f32", "cpuref_pca_fit_f64"):
            f = getattr(L, name)
            f.restype = C.c_int
            f.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                          C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                          C.c_void_p]
        L.cpuref_col_sums_f32.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.cpuref_preprocess_f32.argtypes = [C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float,
                                            C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


def host_threads() -> int:
    """Cores this process may run on (torchrun exports OMP_NUM_THREADS=1: the thread count is set explicitly)."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def set_threads(n: int | None = None) -> int:
    n = n or host_threads()
    lib().cpuref_set_threads(int(n))
    return int(n)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def synth_rows(spec, row0: int, row1: int):
    """Rows [row0, row1) of the spec's matrix as (indptr int64, indices uint32, data float32) — bit-identical to
    single-algebra_b200/synth.py::generate_rows and the device generator."""
    n = int(row1 - row0)
    base = np.ascontiguousarray(spec.base_level, dtype=np.uint8)
    sf = np.ascontiguousarray(spec.sf_offset, dtype=np.int32)
    cdf = np.ascontiguousarray(spec.cdf, dtype=np.uint32)
    ptr = np.empty(n + 1, dtype=np.int64)
    L = lib()
    nnz = L.cpuref_synth(int(spec.seed), int(row0), n, int(spec.ncols), int(spec.n_clusters), _p(base), _p(sf), _p(cdf),
                         _p(ptr), None, None)
    idx = np.empty(nnz, dtype=np.uint32)
    val = np.empty(nnz, dtype=np.float32)
    L.cpuref_synth(int(spec.seed), int(row0), n, int(spec.ncols), int(spec.n_clusters), _p(base), _p(sf), _p(cdf),
                   _p(ptr), _p(idx), _p(val))
    return ptr, idx, val


class CpuPCAResult:
    pass


def pca_fit(indptr, indices, data, nrows, ncols, n_components, omega, mask=None, n_oversamples=10,
            n_power_iterations=7, center=True, want_scores=True):
    """SparsePCA / MaskedSparsePCA fit (+ transform of the fitted rows) with SVDMethod::Random on the host cores.
    `data` float32 or float64 selects the arithmetic type; `omega` is (kept columns) x (rank + n_oversamples)."""
    dt = np.dtype(data.dtype)
    assert dt in (np.dtype(np.float32), np.dtype(np.float64))
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.uint32)
    data = np.ascontiguousarray(data)
    m = None if mask is None else np.ascontiguousarray(np.asarray(mask, dtype=bool).astype(np.uint8))
    n_eff = int(ncols if m is None else m.sum())
    rank = int(min(n_components, nrows, n_eff))
    l = rank + n_oversamples
    omega = np.ascontiguousarray(omega, dtype=dt)
    assert omega.shape == (n_eff, l), (omega.shape, (n_eff, l))
    r = CpuPCAResult()
    r.components = np.empty((rank, n_eff), dtype=dt)
    r.singular_values = np.empty(rank, dtype=np.float64)
    r.explained_variance = np.empty(rank, dtype=np.float64)
    r.mean = np.empty(ncols, dtype=np.float64)
    tv = np.zeros(1, dtype=np.float64)
    r.scores = np.empty((nrows, rank), dtype=dt) if want_scores else None
    tm = np.zeros(8, dtype=np.float64)
    f = lib().cpuref_pca_fit_f32 if dt == np.float32 else lib().cpuref_pca_fit_f64
    rc = f(int(nrows), int(ncols), _p(indptr), _p(indices), _p(data), _p(m), int(n_components), int(n_oversamples),
           int(n_power_iterations), int(bool(center)), _p(omega), _p(r.components), _p(r.singular_values),
           _p(r.explained_variance), _p(r.mean), _p(tv), _p(r.scores), _p(tm))
    if rc != 0:
        raise RuntimeError(f"cpuref_pca_fit failed with code {rc}")
    r.total_var = float(tv[0])
    r.timings = dict(zip(("stats3", "svdlib_means", "products", "qr", "small_svd", "transform", "total"), tm[:7].tolist()))
    return r


def col_sums_f32(indptr, indices, data, nrows, ncols, squared=False):
    out = np.empty(ncols, dtype=np.float64)
    lib().cpuref_col_sums_f32(int(nrows), int(ncols), _p(np.ascontiguousarray(indptr, dtype=np.int64)),
                              _p(np.ascontiguousarray(indices, dtype=np.uint32)),
                              _p(np.ascontiguousarray(data, dtype=np.float32)), int(bool(squared)), _p(out))
    return out


def preprocess_f32(indptr, indices, data, nrows, ncols, target):
    """sum_row -> normalize(ROW, target) -> log1p -> (sum_col, sum_col_squared), in place on `data` (float32)."""
    assert data.dtype == np.float32 and data.flags.c_contiguous and data.flags.writeable
    s = np.empty(ncols, dtype=np.float64)
    q = np.empty(ncols, dtype=np.float64)
    lib().cpuref_preprocess_f32(int(nrows), int(ncols), _p(np.ascontiguousarray(indptr, dtype=np.int64)),
                                _p(np.ascontiguousarray(indices, dtype=np.uint32)), _p(data), float(target), _p(s), _p(q))
    return s, q
