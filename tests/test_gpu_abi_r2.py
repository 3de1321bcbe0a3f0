"""Round-2 additions to the boundary: device CSR <-> CSC transpose, MatrixNonZero counts, value clone / restore,
fit_transform under REFERENCE_COMPAT, sketch width clamped to the operator's dimensions."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import planted_counts
from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_device_transpose_is_scipy_csc(salg, ctx, dt):
    A = planted_counts(700, 300, seed=3, dtype=dt)
    d = salg.CsrMatrix.from_scipy(A, ctx).to_device()
    t = d.transpose()
    assert (t.nrows, t.ncols, t.nnz) == (300, 700, A.nnz)
    off, idx, val = t.download()
    C = sp.csc_matrix(A)
    C.sort_indices()
    assert np.array_equal(off.astype(np.int64), C.indptr)
    assert np.array_equal(idx.astype(np.int64), C.indices)       # rows ascending within a column (stable sort)
    assert np.array_equal(val, C.data)
    # transposing back restores the CSR bit for bit
    b = t.transpose()
    off2, idx2, val2 = b.download()
    assert np.array_equal(off2.astype(np.int64), A.indptr) and np.array_equal(idx2.astype(np.int64), A.indices)
    assert np.array_equal(val2, A.data)


def test_nonzero_row_and_col(salg, ctx):
    """MatrixNonZero::nonzero_row / nonzero_col (src/sparse/csr.rs:23-122): stored entries, explicit zeros included."""
    A = planted_counts(900, 250, seed=5, dtype=np.float32)
    A.data[::17] = 0.0                                            # explicit zeros still count (row_offsets difference)
    x = salg.CsrMatrix(900, 250, A.indptr, A.indices.astype(np.uint64), A.data, ctx)
    assert np.array_equal(x.nonzero_row(), np.diff(A.indptr).astype(np.uint64))
    assert np.array_equal(x.nonzero_col(), np.bincount(A.indices, minlength=250).astype(np.uint64))
    e = salg.CsrMatrix(0, 5, np.zeros(1, np.uint64), np.zeros(0, np.uint64), np.zeros(0, np.float32), ctx)
    assert len(e.nonzero_row()) == 0 and np.array_equal(e.nonzero_col(), np.zeros(5, np.uint64))


def test_values_clone_restore_roundtrip(salg, ctx):
    A = planted_counts(500, 200, seed=9, dtype=np.float32)
    d = salg.CsrMatrix.from_scipy(A, ctx).to_device()
    h = d.clone_values()
    d.preprocess(1e4)
    assert not np.array_equal(d.download_values(), A.data)
    d.restore_values(h)
    assert np.array_equal(d.download_values(), A.data)
    s1, q1 = d.preprocess(1e4)
    v1 = d.download_values()
    d.restore_values(h)
    s2, q2 = d.preprocess(1e4)                                     # repeatable from the same raw counts
    assert np.array_equal(v1, d.download_values())                 # element-wise chain: bit-identical
    assert np.allclose(s1, s2, rtol=1e-5) and np.allclose(q1, q2, rtol=1e-5)   # column sums: f32 atomics reorder
    d.free_values_clone(h)


@pytest.mark.parametrize("masked", [False, True])
def test_fit_transform_follows_transform_mode(salg, ctx, masked):
    """fit_transform = fit + transform (pca/sparse/mod.rs:355-358) also under REFERENCE_COMPAT."""
    A = planted_counts(1200, 300, seed=11, dtype=np.float64)
    mask = np.zeros(300, bool); mask[::3] = True
    b = salg.MaskedSparsePCABuilder().mask(mask.tolist()) if masked else salg.SparsePCABuilder()
    pca = b.n_components(8).svd_method(salg.SVDMethod.Random(6, 4, salg.PowerIterationNormalizer.QR)).build()
    pca.transform_mode = salg.TRANSFORM_REFERENCE_COMPAT
    om = np.random.default_rng(0).standard_normal((100 if masked else 300, 14))
    x = salg.CsrMatrix.from_scipy(A, ctx)
    ft = pca.fit_transform(x, omega=om)
    t = pca.transform(x)
    assert np.array_equal(ft, t)
    ref = O.transform(A, pca.components_, pca.mean_, mask=mask if masked else None, mode=O.REFERENCE_COMPAT)
    assert np.abs(ft - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("masked", [False, True])
def test_sketch_wider_than_operator_is_clamped(salg, ctx, dt, masked):
    """n_eff < n_components + n_oversamples (builder defaults 50 + 10): the sketch is clamped to the operator's dimensions
    instead of running CholeskyQR on a rank-deficient panel; the result is then the exact truncated SVD."""
    ncols = 90 if masked else 55                                   # unmasked: 55 columns < 50 + 10
    A = planted_counts(800, ncols, seed=13, dtype=dt)
    mask = np.zeros(ncols, bool); mask[:40] = True
    n_eff = 40 if masked else ncols
    k = 35 if masked else 50
    p = 25 if masked else 10
    b = salg.MaskedSparsePCABuilder().mask(mask.tolist()) if masked else salg.SparsePCABuilder()
    pca = b.n_components(k).svd_method(salg.SVDMethod.Random(p, 7, salg.PowerIterationNormalizer.QR)).build()
    om = np.random.default_rng(1).standard_normal((n_eff, k + p)).astype(dt)  # as wide as the request; the lead is used
    pca.fit(salg.CsrMatrix.from_scipy(A, ctx), omega=om)
    assert pca.numeric_flags() == 0
    Ak = A.toarray().astype(np.float64)[:, mask] if masked else A.toarray().astype(np.float64)
    Ak = Ak - Ak.mean(axis=0)
    s = np.linalg.svd(Ak, compute_uv=False)[:k]
    tol = 1e-9 if dt == np.float64 else 2e-4
    assert O.rel_err(pca.singular_values_[:k - 5], s[:k - 5]) < tol
