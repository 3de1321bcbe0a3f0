"""Oracle restatement of the SVD engine (un-vendored single-svdlib 1.0.9 — parity unpinned by the
reference) cross-checked against scikit-learn's randomized_svd/svd_flip (which the reference README
credits) and a dense LAPACK SVD; mask compaction against scipy column selection."""
import os

import numpy as np
import scipy.sparse as sp

from conftest import planted_counts
from oracle import oracle as O


def test_randomized_matches_sklearn_range_finder_structure():
    # same Omega, same schedule (QR normaliser, q iterations) => same subspace as scikit-learn's
    # randomized_range_finder + projection, on the explicitly centred dense matrix
    A = planted_counts(600, 150, seed=1)
    rng = np.random.default_rng(3)
    k, p, q = 10, 10, 4
    om = rng.standard_normal((150, k + p))
    u, s, vt = O.randomized_svd(A, k, p, q, om, mean_center=True)
    D = A.toarray()
    D = D - D.mean(axis=0)
    # scikit-learn's algorithm written out with the same Omega
    Q = D @ om
    for _ in range(q):
        Q, _ = np.linalg.qr(Q)
        Q, _ = np.linalg.qr(D.T @ Q)
        Q = D @ Q
    Q, _ = np.linalg.qr(Q)
    B = Q.T @ D
    ub, sb, vtb = np.linalg.svd(B, full_matrices=False)
    assert O.rel_err(s, sb[:k]) < 1e-10
    assert O.largest_principal_angle(vt, vtb[:k]) < 1e-7


def test_randomized_close_to_exact_on_decaying_spectrum():
    A = planted_counts(800, 120, n_clusters=6, seed=2)
    om = np.random.default_rng(0).standard_normal((120, 15))
    u, s, vt = O.randomized_svd(A, 5, 10, 7, om, mean_center=True)
    D = A.toarray()
    D = D - D.mean(axis=0)
    st = np.linalg.svd(D, compute_uv=False)[:5]
    assert O.rel_err(s, st) < 1e-4   # randomized vs exact: approximation error, not parity


def test_normalizers_span_same_subspace():
    # QR / LU / none give the same range up to rounding with the same Omega
    A = planted_counts(500, 100, seed=5)
    om = np.random.default_rng(1).standard_normal((100, 20))
    ref = O.randomized_svd(A, 10, 10, 5, om, normalizer="qr")
    for nz in ("lu",):
        got = O.randomized_svd(A, 10, 10, 5, om, normalizer=nz)
        assert O.rel_err(got[1], ref[1]) < 1e-9
        assert O.largest_principal_angle(got[2], ref[2]) < 1e-6


def test_svd_flip_matches_sklearn():
    from sklearn.utils.extmath import svd_flip
    rng = np.random.default_rng(0)
    u = rng.standard_normal((30, 5))
    vt = rng.standard_normal((5, 12))
    u1, v1 = O.svd_flip_v(u.copy(), vt.copy())
    u2, v2 = svd_flip(u.copy(), vt.copy(), u_based_decision=False)
    assert np.array_equal(u1, u2) and np.array_equal(v1, v2)


def test_mask_compact_equals_scipy_column_selection():
    A = planted_counts(300, 90, seed=7)
    mask = np.random.default_rng(2).random(90) < 0.3
    ip, ix, dv = O.mask_compact(A.indptr, A.indices, A.data, mask)
    B = A[:, np.flatnonzero(mask)].tocsr()
    B.sort_indices()
    assert np.array_equal(ip, B.indptr) and np.array_equal(ix, B.indices) and np.array_equal(dv, B.data)
    # all-false and all-true masks
    ip, ix, dv = O.mask_compact(A.indptr, A.indices, A.data, np.zeros(90, bool))
    assert ip.tolist() == [0] * 301 and len(ix) == 0
    ip, ix, dv = O.mask_compact(A.indptr, A.indices, A.data, np.ones(90, bool))
    assert np.array_equal(ix, A.indices)


def test_fit_semantics_explained_variance_and_mean():
    A = planted_counts(400, 80, seed=9)
    om = np.random.default_rng(4).standard_normal((80, 18))
    r = O.sparse_pca_fit(A, 8, omega=om, n_oversamples=10, n_power_iterations=5)
    n = 400
    assert np.allclose(r.mean, np.asarray(A.mean(axis=0)).ravel())
    assert np.allclose(r.explained_variance, r.singular_values ** 2 / (n - 1))
    D = A.toarray()
    assert np.isclose(r.total_var, D.var(axis=0, ddof=1).sum())
    # masked: mean_ keeps the full length, components live in the compact column space
    mask = np.zeros(80, bool)
    mask[::3] = True
    omm = np.random.default_rng(4).standard_normal((mask.sum(), 18))
    rm = O.sparse_pca_fit(A, 8, omega=omm, mask=mask, n_power_iterations=5)
    assert rm.mean.shape == (80,) and rm.components.shape == (8, mask.sum())
    assert np.isclose(rm.total_var, D[:, mask].var(axis=0, ddof=1).sum())


def test_transform_modes():
    A = planted_counts(200, 40, seed=11)
    om = np.random.default_rng(5).standard_normal((40, 15))
    r = O.sparse_pca_fit(A, 5, omega=om, n_power_iterations=4)
    ex = O.transform(A, r.components, r.mean, mode=O.EXACT)
    D = A.toarray()
    assert np.allclose(ex, (D - r.mean) @ r.components.T)
    # U S is the projection of the exact scores onto range(Q): close, not equal (randomized error)
    assert np.abs(ex - r.u * r.singular_values).max() < 0.05 * np.abs(ex).max()
    # REFERENCE_COMPAT, unmasked: literal loop of pca/sparse/mod.rs:268-282 on two rows
    comp = O.transform(A, r.components, r.mean, mode=O.REFERENCE_COMPAT)
    for row in (0, 17):
        acc = np.zeros(5)
        for c in A.indices:                     # x.col_indices() of the WHOLE matrix
            acc += (D[row, c] - r.mean[c]) * r.components[:, c]
        assert np.allclose(comp[row], acc)
    # masked compat: only stored kept entries (pca/sparse_masked/mod.rs:488-529)
    mask = np.zeros(40, bool)
    mask[1::2] = True
    omm = np.random.default_rng(5).standard_normal((20, 15))
    rm = O.sparse_pca_fit(A, 5, omega=omm, mask=mask, n_power_iterations=4)
    cm = O.transform(A, rm.components, rm.mean, mask=mask, mode=O.REFERENCE_COMPAT)
    kept = np.flatnonzero(mask)
    row = 3
    acc = np.zeros(5)
    for c in A[row].indices:
        if mask[c]:
            acc += (D[row, c] - rm.mean[c]) * rm.components[:, np.searchsorted(kept, c)]
    assert np.allclose(cm[row], acc)


def test_golden_fixture_is_reproducible():
    """tests/golden/pca_small.npz was produced by tests/golden/make_golden.py from this oracle."""
    path = os.path.join(os.path.dirname(__file__), "golden", "pca_small.npz")
    g = np.load(path)
    A = sp.csr_matrix((g["data"], g["indices"], g["indptr"]), shape=tuple(g["shape"]))
    r = O.sparse_pca_fit(A, int(g["k"]), omega=g["omega"], n_oversamples=int(g["p"]),
                         n_power_iterations=int(g["q"]))
    assert O.rel_err(r.singular_values, g["singular_values"]) < 1e-10
    assert O.largest_principal_angle(r.components, g["components"]) < 1e-7
