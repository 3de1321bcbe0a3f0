"""SparsePCA / MaskedSparsePCA through the C ABI against the oracle on the same CSR inputs and the
same host-generated Omega.  Tolerances (north-star): singular values 1e-5 relative in f64, largest
principal angle < 1e-3 rad; mask/index logic bit-exact (tests/test_gpu_csr.py)."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from conftest import planted_counts
from oracle import oracle as O

pytestmark = pytest.mark.gpu
ANGLE_TOL = 1e-3
S_TOL_F64 = 1e-5
S_TOL_F32 = 1e-4     # f32 singular values: not stated by the north-star; f32-vs-f64 measured ~1e-6


def _random(p=10, q=7, norm=None, salg=None):
    return salg.SVDMethod.Random(p, q, salg.PowerIterationNormalizer.QR if norm is None else norm)


def _check_signs(components):
    j = np.argmax(np.abs(components), axis=1)
    assert np.all(components[np.arange(len(j)), j] > 0)


def test_golden_fixture_f64(salg, ctx):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "pca_small.npz"))
    A = sp.csr_matrix((g["data"], g["indices"], g["indptr"]), shape=tuple(g["shape"]))
    x = salg.CsrMatrix.from_scipy(A, ctx)
    pca = salg.SparsePCABuilder().n_components(int(g["k"])).svd_method(_random(int(g["p"]), int(g["q"]), salg=salg)).build()
    scores = pca.fit_transform(x, omega=g["omega"])
    assert O.rel_err(pca.singular_values_, g["singular_values"]) < S_TOL_F64
    assert O.largest_principal_angle(pca.components_, g["components"]) < ANGLE_TOL
    assert O.rel_err(pca.explained_variance_, g["explained_variance"]) < 2 * S_TOL_F64
    assert np.allclose(pca.mean_, g["mean"], rtol=1e-6, atol=0)
    assert abs(pca.total_var_ - float(g["total_var"])) < 1e-6 * float(g["total_var"])
    _check_signs(pca.components_)
    # same signs as the oracle => components and scores agree entrywise
    assert np.abs(pca.components_ - g["components"]).max() < 1e-6
    assert np.abs(scores - g["scores"]).max() < 1e-6 * np.abs(g["scores"]).max()
    # masked fixture
    mp = salg.MaskedSparsePCABuilder().n_components(int(g["k"])).mask(g["mask"].tolist()) \
        .svd_method(_random(int(g["p"]), int(g["q"]), salg=salg)).build()
    mp.fit(x, omega=g["omega_m"])
    assert O.rel_err(mp.singular_values_, g["singular_values_m"]) < S_TOL_F64
    assert O.largest_principal_angle(mp.components_, g["components_m"]) < ANGLE_TOL
    assert abs(mp.total_var_ - float(g["total_var_m"])) < 1e-6 * float(g["total_var_m"])
    assert mp.mean_.shape == (A.shape[1],)


def test_config1_shape_f64_randomized(salg, ctx):
    """BASELINE config 1: 10k x 2k CSR at ~5 %, f64, Random{p=10, q=7, QR}, k=50, centred."""
    spec = salg.synth.make_spec(10_000, 2_000, density=0.05, seed=42)
    ip, ix, dv = salg.synth.generate(spec, dtype=np.float64)
    A = sp.csr_matrix((dv, ix, ip), shape=(10_000, 2_000))
    om = salg.synth.make_omega(2_000, 60, seed=42, dtype=np.float64)
    ref = O.sparse_pca_fit(A, 50, omega=om, n_oversamples=10, n_power_iterations=7)
    pca = salg.SparsePCABuilder().n_components(50).svd_method(_random(salg=salg)).build()
    pca.fit(salg.CsrMatrix.from_scipy(A, ctx), omega=om)
    assert O.rel_err(pca.singular_values_, ref.singular_values) < S_TOL_F64
    assert O.largest_principal_angle(pca.components_, ref.components) < ANGLE_TOL
    assert O.rel_err(pca.explained_variance_, ref.explained_variance) < 2 * S_TOL_F64
    assert np.allclose(pca.explained_variance_ratio(), O.explained_variance_ratio(ref.explained_variance), rtol=1e-4)
    assert np.allclose(pca.cumulative_explained_variance_ratio()[-1], 1.0)
    assert np.allclose(pca.feature_importances(), pca.components_ ** 2)


def test_variance_ratio_against_total_and_noise_variance(salg, ctx):
    """ratio against the total variance and the noise-variance estimate the reference only prints
    (pca/sparse/mod.rs:225-238), against a dense computation."""
    A = planted_counts(1500, 300, seed=9)
    om = salg.synth.make_omega(300, 30, seed=42, dtype=np.float64)
    pca = salg.SparsePCABuilder().n_components(20).svd_method(_random(salg=salg)).build()
    pca.fit(salg.CsrMatrix.from_scipy(A, ctx), omega=om)
    D = A.toarray()
    tv = D.var(axis=0, ddof=1).sum()
    assert abs(pca.total_var_ - tv) < 1e-9 * tv
    assert np.allclose(pca.explained_variance_ratio_total(), pca.explained_variance_ / tv, rtol=1e-9)
    assert pca.explained_variance_ratio_total().sum() < 1.0 < pca.explained_variance_ratio().sum() + 1e-12
    noise = (tv - pca.explained_variance_.sum()) / (min(1500, 300) - 20)
    assert abs(pca.noise_variance() - noise) < 1e-9 * noise


def test_f32_randomized_against_f64_oracle(salg, ctx):
    A = planted_counts(6000, 900, seed=21, dtype=np.float32)
    om = salg.synth.make_omega(900, 40, seed=42, dtype=np.float32)
    ref = O.sparse_pca_fit(A.astype(np.float64), 30, omega=om.astype(np.float64), n_oversamples=10, n_power_iterations=7)
    pca = salg.SparsePCABuilder().n_components(30).svd_method(_random(salg=salg)).build()
    pca.fit(salg.CsrMatrix.from_scipy(A, ctx), omega=om)
    assert pca.components_.dtype == np.float32
    assert O.rel_err(pca.singular_values_, ref.singular_values) < S_TOL_F32
    assert O.largest_principal_angle(pca.components_, ref.components) < ANGLE_TOL
    assert np.allclose(pca.mean_, ref.mean, rtol=1e-4)


@pytest.mark.parametrize("center", [True, False])
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_masked_randomized(salg, ctx, center, dtype):
    A = planted_counts(4000, 1200, seed=22, dtype=dtype)
    mask = salg.synth.make_mask(1200, 300, seed=7)
    om = salg.synth.make_omega(300, 35, seed=42, dtype=dtype)
    ref = O.sparse_pca_fit(A.astype(np.float64), 25, omega=om.astype(np.float64), center=center, mask=mask,
                           n_oversamples=10, n_power_iterations=7)
    pca = salg.MaskedSparsePCABuilder().n_components(25).center(center).mask(mask.tolist()) \
        .svd_method(_random(salg=salg)).build()
    x = salg.CsrMatrix.from_scipy(A, ctx)
    scores = pca.fit_transform(x, omega=om)
    assert pca.components_.shape == (25, 300) and pca.mean_.shape == (1200,)
    assert O.rel_err(pca.singular_values_, ref.singular_values) < (S_TOL_F64 if dtype == np.float64 else S_TOL_F32)
    assert O.largest_principal_angle(pca.components_, ref.components) < ANGLE_TOL
    assert np.allclose(pca.mean_, ref.mean, rtol=1e-4, atol=1e-12)
    assert abs(pca.total_var_ - ref.total_var) < 1e-4 * ref.total_var
    # fit_transform == transform(x) == exact projection with OUR components
    ex = O.transform(A, pca.components_, pca.mean_, center=center, mask=mask, mode=O.EXACT)
    tol = 1e-9 if dtype == np.float64 else 2e-4
    assert np.abs(scores - ex).max() < tol * np.abs(ex).max()
    t = pca.transform(x)
    assert np.abs(t - ex).max() < tol * np.abs(ex).max()
    # the reference's own loop semantics (non-zero-only centring)
    pca.transform_mode = salg.TRANSFORM_REFERENCE_COMPAT
    tc = pca.transform(x)
    rc = O.transform(A, pca.components_, pca.mean_, center=center, mask=mask, mode=O.REFERENCE_COMPAT)
    assert np.abs(tc - rc).max() < tol * max(np.abs(rc).max(), 1.0)


def test_unmasked_transform_modes(salg, ctx):
    A = planted_counts(1500, 250, seed=23)
    om = salg.synth.make_omega(250, 20, seed=1, dtype=np.float64)
    pca = salg.SparsePCABuilder().n_components(10).svd_method(_random(10, 5, salg=salg)).build()
    x = salg.CsrMatrix.from_scipy(A, ctx)
    pca.fit(x, omega=om)
    B = planted_counts(333, 250, seed=24)        # transform rows that were not in the fit
    xb = salg.CsrMatrix.from_scipy(B, ctx)
    ex = O.transform(B, pca.components_, pca.mean_, mode=O.EXACT)
    assert np.abs(pca.transform(xb) - ex).max() < 1e-9 * np.abs(ex).max()
    pca.transform_mode = salg.TRANSFORM_REFERENCE_COMPAT
    rc = O.transform(B, pca.components_, pca.mean_, mode=O.REFERENCE_COMPAT)   # cnt-weighted
    assert np.abs(pca.transform(xb) - rc).max() < 1e-9 * np.abs(rc).max()


def test_normalizers_agree(salg, ctx):
    A = planted_counts(3000, 400, seed=25)
    om = salg.synth.make_omega(400, 30, seed=3, dtype=np.float64)
    x = salg.CsrMatrix.from_scipy(A, ctx)
    res = []
    for nz in (salg.PowerIterationNormalizer.QR, salg.PowerIterationNormalizer.LU,
               salg.PowerIterationNormalizer.NoNormalization):
        pca = salg.SparsePCABuilder().n_components(20).svd_method(_random(10, 7, nz, salg=salg)).build()
        pca.fit(x, omega=om)
        res.append(pca)
    ref = O.sparse_pca_fit(A, 20, omega=om, n_oversamples=10, n_power_iterations=7, normalizer="lu")
    for r in res:
        assert O.rel_err(r.singular_values_, ref.singular_values) < S_TOL_F64
        assert O.largest_principal_angle(r.components_, ref.components) < ANGLE_TOL


def test_device_omega_is_statistically_consistent(salg, ctx):
    # omega == NULL: the library draws its own test matrix; only the well-separated leading part is comparable
    A = planted_counts(4000, 500, n_clusters=6, seed=26)
    pca = salg.SparsePCABuilder().n_components(5).svd_method(_random(10, 7, salg=salg)).build()
    pca.fit(salg.CsrMatrix.from_scipy(A, ctx))
    D = A.toarray()
    D -= D.mean(axis=0)
    st = np.linalg.svd(D, compute_uv=False)[:5]
    assert O.rel_err(pca.singular_values_, st) < 1e-3


@pytest.mark.parametrize("dtype,stol", [(np.float64, 1e-5), (np.float32, 1e-4)])
def test_lanczos_uncentred_operator(salg, ctx, dtype, stol):
    """SVDMethod::Lanczos runs on the UNCENTRED matrix even with center=true."""
    A = planted_counts(3000, 420, seed=27, dtype=dtype)
    x = salg.CsrMatrix.from_scipy(A, ctx)
    pca = salg.SparsePCABuilder().n_components(20).svd_method(salg.SVDMethod.Lanczos).build()
    scores = pca.fit_transform(x)
    u, s, vt = O.truncated_svd_truth(A.astype(np.float64), 20)
    assert O.rel_err(pca.singular_values_, s) < stol
    assert O.largest_principal_angle(pca.components_, vt) < ANGLE_TOL
    assert np.allclose(pca.mean_, np.asarray(A.mean(axis=0)).ravel(), rtol=1e-4)
    _check_signs(pca.components_)
    ex = O.transform(A, pca.components_, pca.mean_, mode=O.EXACT)
    assert np.abs(scores - ex).max() < (1e-9 if dtype == np.float64 else 2e-4) * np.abs(ex).max()


def test_lanczos_masked_small_matrix_runs_to_full_dimension(salg, ctx):
    A = planted_counts(90, 60, seed=28)
    mask = np.zeros(60, bool)
    mask[::2] = True
    pca = salg.MaskedSparsePCABuilder().n_components(10).mask(mask.tolist()).build()      # default method: Lanczos
    pca.fit(salg.CsrMatrix.from_scipy(A, ctx))
    u, s, vt = O.truncated_svd_truth(A[:, np.flatnonzero(mask)].astype(np.float64), 10)
    assert O.rel_err(pca.singular_values_, s) < 1e-8
    assert O.largest_principal_angle(pca.components_, vt) < ANGLE_TOL


def test_errors_match_reference_behaviour(salg, ctx):
    A = planted_counts(200, 50, seed=29)
    x = salg.CsrMatrix.from_scipy(A, ctx)
    mp = salg.MaskedSparsePCABuilder().n_components(5).mask([True] * 49).svd_method(_random(salg=salg)).build()
    with pytest.raises(salg.SalgError) as e:
        mp.fit(x)
    assert e.value.code == 2
    assert str(e.value) == "The mask vector length and the number of features (columns) have to be the same!"
    p = salg.SparsePCABuilder().build()
    with pytest.raises(salg.SalgError) as e:
        p.transform(x)
    assert e.value.code == 3 and str(e.value) == "Must be fitted before transform!"
    big = salg.SparsePCABuilder().n_components(60).svd_method(_random(10, 2, salg=salg)).build()
    with pytest.raises(salg.SalgError) as e:
        big.fit(salg.CsrMatrix.from_scipy(planted_counts(300, 200, seed=30), ctx))
    assert e.value.code == 7
    # builder defaults (pca/sparse/mod.rs:388-403)
    assert (p.n_components, p.alpha, p.tolerance, p.random_seed, p.center, p.verbose) == (50, 1.0, 1e-6, 42, True, False)
    assert p.svdmethod == salg.SVDMethod.Lanczos


def test_masked_statistics_integer_accumulators_fall_back(salg, ctx):
    """The masked statistics pass accumulates raw counts in integer shared-memory atomics after probing a prefix of the
    values; a non-integral (or >= 65536, or negative) value past the probed prefix must send it back to f32 accumulators."""
    A = planted_counts(3000, 700, seed=33, dtype=np.float32).tocsr()
    assert A.nnz > 70_000
    for late in (2.5, 70000.0, -3.0):
        B = A.copy()
        B.data[-5] = late
        mask = salg.synth.make_mask(700, 200, seed=7)
        om = salg.synth.make_omega(200, 30, seed=42, dtype=np.float32)
        pca = salg.MaskedSparsePCABuilder().n_components(20).mask(mask.tolist()).svd_method(_random(salg=salg)).build()
        pca.fit(salg.CsrMatrix.from_scipy(B, ctx), omega=om)
        ref = O.sparse_pca_fit(B.astype(np.float64), 20, omega=om.astype(np.float64), mask=mask, n_oversamples=10,
                               n_power_iterations=7)
        assert np.allclose(pca.mean_, ref.mean, rtol=1e-5, atol=1e-7)
        assert abs(pca.total_var_ - ref.total_var) < 1e-4 * ref.total_var
        assert O.largest_principal_angle(pca.components_, ref.components) < ANGLE_TOL


def _masked_fit_against_oracle(salg, ctx, A, mask, k, l_extra=10, q=7):
    n_eff = int(mask.sum())
    om = salg.synth.make_omega(n_eff, k + l_extra, seed=42, dtype=np.float32)
    pca = salg.MaskedSparsePCABuilder().n_components(k).mask(mask.tolist()).svd_method(
        _random(l_extra, q, salg=salg)).build()
    scores = pca.fit_transform(salg.CsrMatrix.from_scipy(A, ctx), omega=om)
    ref = O.sparse_pca_fit(A.astype(np.float64), k, omega=om.astype(np.float64), mask=mask, n_oversamples=l_extra,
                           n_power_iterations=q)
    return pca, scores, ref


def test_fused_compaction_slot_overflow_falls_back(salg, ctx):
    """The statistics pass writes the kept entries of a row into a slot sized from the GLOBAL kept fraction; a row whose
    entries all sit in kept columns overflows its slot and the fit must fall back to the separate compaction pass."""
    A = planted_counts(2500, 800, seed=17, dtype=np.float32).tolil()
    mask = np.zeros(800, dtype=bool)
    kept = np.sort(np.random.default_rng(4).choice(800, 40, replace=False))      # 5 % of the columns -> slots = len / 4
    mask[kept] = True
    for r in (3, 1200, 2499):                        # rows living entirely inside the kept columns
        A.rows[r] = kept.tolist()
        A.data[r] = [float(1 + (j % 5)) for j in range(len(kept))]
    A = A.tocsr().astype(np.float32)
    A.sort_indices()
    pca, scores, ref = _masked_fit_against_oracle(salg, ctx, A, mask, 12)
    assert O.rel_err(pca.singular_values_, ref.singular_values) < S_TOL_F32
    assert O.largest_principal_angle(pca.components_, ref.components) < ANGLE_TOL
    assert np.allclose(pca.mean_, ref.mean, rtol=1e-5, atol=1e-7)
    ex = O.transform(A, pca.components_, pca.mean_, mask=mask, mode=O.EXACT)
    assert np.abs(scores - ex).max() < 3e-4 * np.abs(ex).max()


@pytest.mark.parametrize("nrows", [70, 129, 1000])
def test_masked_f32_edge_shapes(salg, ctx, nrows):
    """Fewer rows than one 128-row tile, one row over, empty rows, rows without any kept entry, a mask keeping fewer
    columns than one 64-column tile: the fused statistics / tile / Gram passes must agree with the oracle."""
    A = planted_counts(nrows, 300, density=0.15, seed=nrows, dtype=np.float32).tolil()
    A.rows[0], A.data[0] = [], []                    # empty rows at both ends
    A.rows[nrows - 1], A.data[nrows - 1] = [], []
    A = A.tocsr().astype(np.float32)
    mask = np.zeros(300, dtype=bool)
    mask[np.random.default_rng(2).choice(300, 40, replace=False)] = True
    pca, scores, ref = _masked_fit_against_oracle(salg, ctx, A, mask, 8, l_extra=6, q=4)
    assert pca.components_.shape == (8, 40)
    assert O.rel_err(pca.singular_values_, ref.singular_values) < S_TOL_F32
    assert O.largest_principal_angle(pca.components_, ref.components) < ANGLE_TOL
    ex = O.transform(A, pca.components_, pca.mean_, mask=mask, mode=O.EXACT)
    assert scores.shape == (nrows, 8)
    assert np.abs(scores - ex).max() < 3e-4 * np.abs(ex).max()
    assert np.abs(scores[0] - ex[0]).max() < 3e-4 * np.abs(ex).max()     # an empty row projects to -mu V


@pytest.mark.parametrize("q,p_over", [(0, 10), (1, 0), (2, 5)])
def test_f32_power_iteration_counts(salg, ctx, q, p_over):
    """n_power_iterations = 0 skips the fused tall-panel path entirely, 1 runs it once; n_oversamples = 0 makes l = k.
    Same Omega on both sides, so the result must track the oracle for every count."""
    A = planted_counts(3000, 500, seed=51, dtype=np.float32)
    k = 12
    om = salg.synth.make_omega(500, k + p_over, seed=42, dtype=np.float32)
    pca = salg.SparsePCABuilder().n_components(k).svd_method(_random(p_over, q, salg=salg)).build()
    pca.fit(salg.CsrMatrix.from_scipy(A, ctx), omega=om)
    ref = O.sparse_pca_fit(A.astype(np.float64), k, omega=om.astype(np.float64), n_oversamples=p_over, n_power_iterations=q)
    assert O.rel_err(pca.singular_values_, ref.singular_values) < S_TOL_F32
    # without power iterations the trailing components of a sketch are ill-determined (both sides compute the same
    # sketch; their angle is governed by the f32 rounding of tiny gaps), so compare the leading half there
    lead = k if q > 0 else k // 2
    assert O.largest_principal_angle(pca.components_[:lead], ref.components[:lead]) < (ANGLE_TOL if q > 0 else 5e-3)


@pytest.mark.parametrize("dtype,ncols", [(np.float64, 20000), (np.float32, 50000)])
def test_masked_fit_through_the_column_tiled_statistics(salg, ctx, dtype, ncols):
    """Wide masked fits whose accumulators do not fit one shared-memory tile (f64 beyond ~12.7k columns, f32 beyond ~45k):
    the column-tiled statistics kernel sweeps each row block once per tile and carries the kept-entry count of a row across
    the tiles; the compaction built from those counts feeds the fit, which must match the oracle."""
    A = planted_counts(1000, ncols, density=0.02, seed=ncols, dtype=dtype)
    n_keep = 400
    mask = salg.synth.make_mask(ncols, n_keep, seed=7)
    x = salg.CsrMatrix.from_scipy(A, ctx)
    om = salg.synth.make_omega(n_keep, 20, seed=42, dtype=dtype)
    ref = O.sparse_pca_fit(A.astype(np.float64), 10, omega=om.astype(np.float64), mask=mask, n_oversamples=10,
                           n_power_iterations=7)
    pca = salg.MaskedSparsePCABuilder().n_components(10).mask(mask.tolist()).svd_method(_random(salg=salg)).build()
    pca.fit(x, omega=om)
    assert pca.components_.shape == (10, n_keep) and pca.mean_.shape == (ncols,)
    assert np.allclose(pca.mean_, ref.mean, rtol=(1e-6 if dtype == np.float64 else 1e-4), atol=0)
    assert abs(pca.total_var_ - ref.total_var) < (1e-6 if dtype == np.float64 else 1e-4) * ref.total_var
    if dtype == np.float64:
        assert O.rel_err(pca.singular_values_, ref.singular_values) < S_TOL_F64
        assert O.largest_principal_angle(pca.components_, ref.components) < ANGLE_TOL
    else:
        # singular values only: the trailing components of this thin matrix sit in the noise floor, where an f32 subspace
        # is gap-limited whatever the compaction does
        assert O.rel_err(pca.singular_values_, ref.singular_values) < 1e-3
