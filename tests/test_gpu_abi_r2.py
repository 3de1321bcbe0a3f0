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
    assert np.allclose(ft, t, rtol=1e-11, atol=1e-9 * np.abs(t).max())      # (atomic accumulation order differs run to run)
    ref = O.transform(A, pca.components_, pca.mean_, mask=mask if masked else None, mode=O.REFERENCE_COMPAT)
    assert np.abs(ft - ref).max() < 1e-9 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("masked", [False, True])
def test_sketch_wider_than_operator_is_clamped(salg, ctx, dt, masked):
    """n_eff < n_components + n_oversamples (builder defaults 50 + 10): the sketch is clamped to the operator's dimensions
    instead of running CholeskyQR on a rank-deficient panel; the result is then the exact truncated SVD."""
    ncols = 90 if masked else 55                                   # unmasked: 55 columns < 50 + 10
    A = planted_counts(800, ncols, density=0.5, seed=13, dtype=dt)   # dense enough that no column is empty (full rank)
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


def _host_csr(salg, ctx, A):
    return salg.CsrMatrix(A.shape[0], A.shape[1], A.indptr.astype(np.int64).view(np.uint64), A.indices.astype(np.int32),
                          A.data.astype(np.float32), ctx)


@pytest.mark.parametrize("chunk", [None, 4000, 300])
def test_streamed_host_fit_matches_resident_fit(salg, ctx, chunk, monkeypatch):
    """salg_pca_fit_host_f32: the matrix streams through the statistics + compaction pass in row chunks;
    the model must equal the fit of the resident upload (same kernels, same order of the integer sums)."""
    A = planted_counts(3000, 700, seed=21, dtype=np.float32)
    A[17] = 0; A[18] = 0                                       # empty rows inside a chunk
    A.eliminate_zeros()
    mask = np.zeros(700, bool); mask[::4] = True
    om = np.random.default_rng(2).standard_normal((175, 30)).astype(np.float32)
    def build():
        return salg.MaskedSparsePCABuilder().mask(mask.tolist()).n_components(20).svd_method(
            salg.SVDMethod.Random(10, 5, salg.PowerIterationNormalizer.QR)).build()
    ref = build()
    dev = _host_csr(salg, ctx, A).to_device()
    sc_ref = ref.fit_transform(dev, omega=om)
    if chunk:
        monkeypatch.setenv("SALG_STREAM_CHUNK", str(chunk))
    p = build()
    sc = p.fit_transform(_host_csr(salg, ctx, A), omega=om)      # host matrix -> streamed
    assert np.array_equal(p.mean_, ref.mean_) and p.total_var_ == ref.total_var_      # integer sums: exact
    assert O.rel_err(p.singular_values_, ref.singular_values_) < 1e-5
    assert O.largest_principal_angle(p.components_, ref.components_) < 1e-3
    assert np.abs(sc - sc_ref).max() < 2e-3 * np.abs(sc_ref).max()
    oracle = O.sparse_pca_fit(sp.csr_matrix(A, dtype=np.float64), 20, omega=om.astype(np.float64), mask=mask, n_oversamples=10,
                              n_power_iterations=5)
    assert O.rel_err(p.singular_values_, oracle.singular_values) < 1e-4


def test_streamed_host_fit_falls_back_on_non_count_values(salg, ctx, monkeypatch):
    """Values that stop being raw counts after the probed prefix (or never were) still give the right model: f32
    accumulators from the start, or a redo from a resident upload."""
    A = planted_counts(2500, 400, seed=23, dtype=np.float32)
    A.data[A.nnz // 2:] += 0.25                                # integers first, fractions in later chunks
    mask = np.zeros(400, bool); mask[1::3] = True
    om = np.random.default_rng(3).standard_normal((int(mask.sum()), 24)).astype(np.float32)
    monkeypatch.setenv("SALG_STREAM_CHUNK", "20000")
    p = salg.MaskedSparsePCABuilder().mask(mask.tolist()).n_components(16).svd_method(
        salg.SVDMethod.Random(8, 5, salg.PowerIterationNormalizer.QR)).build()
    p.fit(_host_csr(salg, ctx, A), omega=om)
    oracle = O.sparse_pca_fit(sp.csr_matrix(A, dtype=np.float64), 16, omega=om.astype(np.float64), mask=mask, n_oversamples=8,
                              n_power_iterations=5)
    assert O.rel_err(p.singular_values_, oracle.singular_values) < 1e-4
    assert np.allclose(p.mean_, oracle.mean, rtol=1e-5, atol=1e-7)


def test_streamed_host_fit_rejects_malformed_rows(salg, ctx):
    A = planted_counts(600, 200, seed=25, dtype=np.float32)
    mask = np.ones(200, bool)
    x = _host_csr(salg, ctx, A)
    bad = x.col_indices.copy()
    s0, s1 = int(A.indptr[5]), int(A.indptr[6])
    assert s1 - s0 >= 2
    bad[s0], bad[s0 + 1] = bad[s0 + 1], bad[s0]                  # not strictly increasing inside row 5
    x.col_indices = bad
    p = salg.MaskedSparsePCABuilder().mask(mask.tolist()).n_components(5).svd_method(
        salg.SVDMethod.Random(5, 2, salg.PowerIterationNormalizer.QR)).build()
    with pytest.raises(salg.SalgError, match="strictly increasing"):
        p.fit(x)
