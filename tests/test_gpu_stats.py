"""MatrixSum / Normalize / Log1P on the device against the oracle and the reference's KATs.
Tolerances are the north-star's: column statistics 1e-6 relative in f64, 1e-4 in f32."""
import json
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import planted_counts
from oracle import oracle as O

pytestmark = pytest.mark.gpu
KAT = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "kat.json")))
TOL = {np.float32: 1e-4, np.float64: 1e-6}


def _rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-30)) if len(b) else 0.0


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("shape", [(2000, 300), (700, 30000), (300, 70000)])
def test_sums_match_oracle(salg, ctx, dtype, shape):
    # 30000 / 70000 columns exercise the column-tiled kernel (f64: > 12.8k, f32: > 25.6k columns)
    A = planted_counts(shape[0], shape[1], density=0.03, seed=shape[1], dtype=dtype)
    A.data = (A.data + np.random.default_rng(0).random(A.nnz)).astype(dtype)   # non-integer values
    d = salg.CsrMatrix.from_scipy(A, ctx).to_device()
    s, q = d.sum_col_and_squared()
    nz = np.asarray(O.sum_col(A.indptr, A.indices, A.data, A.shape[1])) != 0
    assert _rel(s[nz], O.sum_col(A.indptr, A.indices, A.data, A.shape[1])[nz]) < TOL[dtype]
    assert _rel(q[nz], O.sum_col_squared(A.indptr, A.indices, A.data, A.shape[1])[nz]) < TOL[dtype]
    assert np.all(s[~nz] == 0) and np.all(q[~nz] == 0)
    assert _rel(d.sum_row(), O.sum_row(A.indptr, A.indices, A.data, A.shape[0])) < TOL[dtype]
    assert s.dtype == dtype


def test_kat_s1_exact(salg, ctx):
    k = KAT["KAT-S1"]
    m = salg.CsrMatrix.from_scipy(sp.csr_matrix(np.array(k["dense"])), ctx)
    assert m.sum_col().tolist() == k["col_sums"]
    assert m.sum_row().tolist() == k["row_sums"]
    assert m.sum_col_squared().tolist() == [17.0, 9.0, 29.0]


def test_col_stats_counts_and_variance(salg, ctx):
    A = planted_counts(900, 400, seed=5)
    d = salg.CsrMatrix.from_scipy(A, ctx).to_device()
    s, q, cnt, var = d.col_stats()
    assert np.array_equal(cnt, np.diff(A.tocsc().indptr).astype(np.float64))
    n = 900.0
    ref = (q / n - (s / n) ** 2) * n / (n - 1)      # MatrixVariance::var_col, src/sparse/csr.rs:649-657
    assert np.allclose(var, ref, rtol=1e-12, atol=1e-15)
    assert np.allclose(var, A.toarray().var(axis=0, ddof=1), rtol=1e-9, atol=1e-12)


def test_kat_n1_normalize(salg, ctx):
    k = KAT["KAT-N1"]
    A = sp.coo_matrix((k["vals"], (k["rows"], k["cols"])), shape=(3, 3)).tocsr()
    m = salg.CsrMatrix.from_scipy(A, ctx)
    m.normalize(np.array(k["column"]["sums"]), k["column"]["target"], salg.Direction.COLUMN)
    assert np.max(np.abs(m.values - np.array(k["column"]["expected"]))) < k["tol"]
    m = salg.CsrMatrix.from_scipy(A, ctx)
    m.normalize(np.array(k["row"]["sums"]), k["row"]["target"], salg.Direction.ROW)
    assert np.max(np.abs(m.values - np.array(k["row"]["expected"]))) < k["tol"]


def test_kat_n2_normalize_sums_to_target(salg, ctx):
    k = KAT["KAT-N2"]
    A = sp.csr_matrix(np.array(k["dense"]))
    m = salg.CsrMatrix.from_scipy(A, ctx)
    m.normalize(np.array(k["col_sums"]), k["target"], salg.Direction.COLUMN)
    assert np.max(np.abs(m.sum_col() - k["target"])) < k["tol"]
    m = salg.CsrMatrix.from_scipy(A, ctx)
    m.normalize(np.array(k["row_sums"]), k["target"], salg.Direction.ROW)
    assert np.max(np.abs(m.sum_row() - k["target"])) < k["tol"]


@pytest.mark.parametrize("dtype,udtype", [(np.float32, np.float32), (np.float64, np.float64), (np.float32, np.float64)])
@pytest.mark.parametrize("direction", [0, 1])
def test_normalize_matches_oracle(salg, ctx, dtype, udtype, direction):
    A = planted_counts(800, 350, seed=6, dtype=dtype)
    sums = (O.sum_row(A.indptr, A.indices, A.data, 800) if direction == 0
            else O.sum_col(A.indptr, A.indices, A.data, 350)).astype(udtype)
    sums[::7] = 0            # zero / negative sums: entries stay untouched (csr.rs:1041,1055)
    sums[3::11] = -2
    m = salg.CsrMatrix.from_scipy(A, ctx)
    m.normalize(sums, 1e4, direction)
    ref = O.normalize(A.indptr, A.indices, A.data, sums, 1e4, direction)
    assert m.values.dtype == dtype
    # one rounding of a product in U then a cast: bit-exact unless the division differs by an ulp
    assert _rel(m.values, ref) < (2e-7 if dtype == np.float32 else 4e-16)


def test_normalize_short_sums_is_an_error_not_a_panic(salg, ctx):
    A = planted_counts(40, 20, seed=7)
    m = salg.CsrMatrix.from_scipy(A, ctx)
    with pytest.raises(salg.SalgError) as e:
        m.normalize(np.ones(10), 1.0, salg.Direction.ROW)
    assert e.value.code == 1


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_log1p_matches_oracle(salg, ctx, dtype):
    A = planted_counts(600, 200, seed=8, dtype=dtype)
    A.data = (A.data * np.random.default_rng(1).random(A.nnz) * 3).astype(dtype)
    m = salg.CsrMatrix.from_scipy(A, ctx)
    m.log1p_normalize()
    ref = O.log1p_normalize(A.data)
    # ln(fl(1+x)) with the device's log: within 2 ulp of the host libm
    assert _rel(m.values, ref) < (3e-7 if dtype == np.float32 else 5e-16)
    k = KAT["KAT-L1"]
    z = salg.CsrMatrix(2, 2, np.array([0, 1, 2], np.uint64), np.array([0, 1], np.uint64), np.array([0.0, 0.0]), ctx)
    z.log1p_normalize()
    assert np.max(np.abs(z.values)) < k["tol"]


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_fused_preprocess_matches_reference_chain(salg, ctx, dtype):
    A = planted_counts(1000, 260, seed=9, dtype=dtype)
    A = sp.vstack([A, sp.csr_matrix((3, 260), dtype=dtype)]).tocsr()    # trailing empty rows
    d = salg.CsrMatrix.from_scipy(A, ctx).to_device()
    s, q = d.preprocess(1e4)
    rs = O.sum_row(A.indptr, A.indices, A.data, A.shape[0])
    v = O.normalize(A.indptr, A.indices, A.data, rs, dtype(1e4), O.ROW)
    v = O.log1p_normalize(v)
    got = d.download_values()
    assert _rel(got, v) < (1e-6 if dtype == np.float32 else 1e-14)
    assert _rel(s, O.sum_col(A.indptr, A.indices, v, 260)) < TOL[dtype]
    assert _rel(q, O.sum_col_squared(A.indptr, A.indices, v, 260)) < TOL[dtype]


# ---- CSC twins (src/sparse/csc.rs): the reference's CSC known-answer tests run on a CscMatrix ----------
def test_csc_kat_s1_exact(salg, ctx):
    k = KAT["KAT-S1"]                                    # csc.rs:1124-1152
    m = salg.CscMatrix.from_scipy(sp.csc_matrix(np.array(k["dense"])), ctx)
    assert m.sum_col().tolist() == k["col_sums"]
    assert m.sum_row().tolist() == k["row_sums"]
    assert m.sum_col_squared().tolist() == [17.0, 9.0, 29.0]


def test_csc_kat_n2_normalize_sums_to_target(salg, ctx):
    k = KAT["KAT-N2"]                                    # csc.rs:1257-1301
    D = np.array(k["dense"], dtype=np.float64)
    for direction, sums, axis in ((salg.Direction.COLUMN, k["col_sums"], 0), (salg.Direction.ROW, k["row_sums"], 1)):
        m = salg.CscMatrix.from_scipy(sp.csc_matrix(D), ctx)
        m.normalize(np.array(sums, dtype=np.float64), float(k["target"]), direction)
        back = sp.csc_matrix((m.values, m.row_indices.astype(np.int64), m.col_offsets.astype(np.int64)), shape=D.shape).toarray()
        got = back.sum(axis=axis)
        want = np.where(np.array(sums) > 0, k["target"], 0.0)
        assert np.abs(got - want).max() < k["tol"]


def test_csc_kat_l1_log1p(salg, ctx):
    k = KAT["KAT-L1"]                                    # csc.rs:1304-1314
    v = np.array(k["vals"], dtype=np.float64)
    m = salg.CscMatrix(len(v), 1, [0, len(v)], np.arange(len(v)), v, ctx)
    m.log1p_normalize()
    assert np.abs(m.values - np.log(1.0 + v)).max() < k["tol"]


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_csc_twins_match_csr(salg, ctx, dtype):
    A = planted_counts(700, 300, seed=12, dtype=dtype)
    c, r = salg.CscMatrix.from_scipy(A, ctx), salg.CsrMatrix.from_scipy(A, ctx)
    tol = 1e-12 if dtype == np.float64 else 1e-6
    assert np.allclose(c.sum_col(), r.sum_col(), rtol=tol)
    assert np.allclose(c.sum_row(), r.sum_row(), rtol=tol)
    assert np.allclose(c.sum_col_squared(), r.sum_col_squared(), rtol=tol)
    for direction, n in ((salg.Direction.ROW, 700), (salg.Direction.COLUMN, 300)):
        sums = np.random.default_rng(3).uniform(0.0, 5.0, n).astype(dtype)
        sums[::7] = 0.0                                  # scale 0: entries untouched (csc.rs:690-697)
        c2, r2 = salg.CscMatrix.from_scipy(A, ctx), salg.CsrMatrix.from_scipy(A, ctx)
        c2.normalize(sums, 100.0, direction)
        r2.normalize(sums, 100.0, direction)
        Bc = sp.csc_matrix((c2.values, c2.row_indices.astype(np.int64), c2.col_offsets.astype(np.int64)), shape=A.shape)
        Br = sp.csr_matrix((r2.values, r2.col_indices.astype(np.int64), r2.row_offsets.astype(np.int64)), shape=A.shape)
        assert abs(Bc - Br).max() == 0.0                 # same arithmetic on both layouts: bit-identical values


# ---- batch boundaries of the row streams (whole batches of 32 x U entries without per-load predicates, clamped tail batch) ------
def _ragged(lengths, ncols, dtype, seed):
    rng = np.random.default_rng(seed)
    indptr = np.zeros(len(lengths) + 1, np.int64)
    indptr[1:] = np.cumsum(lengths)
    idx = np.concatenate([np.sort(rng.choice(ncols, n, replace=False)) for n in lengths] or [np.zeros(0, np.int64)])
    val = (rng.integers(1, 9, indptr[-1]) + rng.random(indptr[-1])).astype(dtype)
    return sp.csr_matrix((val, idx.astype(np.int64), indptr), shape=(len(lengths), ncols))


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("ncols", [6000, 70000])
def test_row_streams_at_batch_boundaries(salg, ctx, dtype, ncols):
    """Row lengths around every batch size in use (32 x 8 = 256, 32 x 16 = 512, 256-thread CTAs x 8 = 2048) plus empty and
    one-entry rows; 70000 columns puts the column statistics on the tiled kernel with whole batches that straddle tile
    borders."""
    lengths = [0, 1, 2, 31, 32, 33, 255, 256, 257, 0, 511, 512, 513, 767, 768, 1023, 1024, 1025, 2047, 2048, 2049, 4100, 5999, 0]
    A = _ragged(lengths, ncols, dtype, seed=ncols)
    m = salg.CsrMatrix.from_scipy(A, ctx)
    d = m.to_device()
    s, q = d.sum_col_and_squared()
    assert np.allclose(s, O.sum_col(A.indptr, A.indices, A.data, ncols), rtol=TOL[dtype], atol=0)
    assert np.allclose(q, O.sum_col_squared(A.indptr, A.indices, A.data, ncols), rtol=TOL[dtype], atol=0)
    rs = d.sum_row()
    assert np.allclose(rs, O.sum_row(A.indptr, A.indices, A.data, A.shape[0]), rtol=TOL[dtype], atol=0)
    # transposed SpMV-style check of the Lanczos product kernels is covered by test_gpu_pca; here the in-place streams
    ref = O.normalize(A.indptr, A.indices, A.data, rs, dtype(1e4), O.ROW)
    m.normalize(rs, dtype(1e4), salg.Direction.ROW)
    assert np.allclose(m.values, ref, rtol=(2e-6 if dtype == np.float32 else 1e-14), atol=0)
    m.log1p_normalize()
    assert np.allclose(m.values, O.log1p_normalize(ref), rtol=(1e-6 if dtype == np.float32 else 1e-14), atol=0)
    # fused chain on a fresh copy
    d2 = salg.CsrMatrix.from_scipy(A, ctx).to_device()
    d2.preprocess(1e4)
    assert np.allclose(d2.download_values(), O.log1p_normalize(ref), rtol=(2e-6 if dtype == np.float32 else 1e-14), atol=0)
    # column normalisation (flat kernel, 4 entries per thread, tail of < 4 entries)
    m3 = salg.CsrMatrix.from_scipy(A, ctx)
    cs = m3.sum_col()
    m3.normalize(cs, dtype(1.0), salg.Direction.COLUMN)
    ref3 = O.normalize(A.indptr, A.indices, A.data, cs, dtype(1.0), O.COLUMN)
    assert np.allclose(m3.values, ref3, rtol=(2e-6 if dtype == np.float32 else 1e-14), atol=0)
