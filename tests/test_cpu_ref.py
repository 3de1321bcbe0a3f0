"""The C++/OpenMP CPU restatement (oracle/cpu_ref.cpp) against the numpy oracle and the host generators.
CPU-only: this is checker-vs-checker, so that the full-size GPU parity tests and the bench's reference arm stand on
a CPU path that is itself pinned to oracle/oracle.py (which is pinned to the reference's KATs / scikit-learn)."""
import numpy as np
import pytest
import scipy.sparse as sp

import single_algebra_b200 as s
from oracle import cpu_ref as R
from oracle import oracle as O


@pytest.fixture(scope="module")
def small():
    spec = s.synth.make_spec(3000, 900, density=0.07, seed=42)
    ip, ix, dv = s.synth.generate_rows(spec, 0, 3000, dtype=np.float32)
    return spec, ip, ix, dv


def test_host_generator_matches_numpy_generator(small):
    spec, ip, ix, dv = small
    ptr, idx, val = R.synth_rows(spec, 0, 3000)
    assert np.array_equal(ptr, ip) and np.array_equal(idx.astype(np.int64), ix) and np.array_equal(val, dv)
    ptr2, idx2, val2 = R.synth_rows(spec, 1000, 1500)        # any row range regenerates bit-identically
    assert np.array_equal(ptr2, ip[1000:1501] - ip[1000])
    assert np.array_equal(val2, dv[ip[1000]:ip[1500]])


@pytest.mark.parametrize("masked", [False, True])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_cpu_ref_fit_matches_numpy_oracle(small, masked, dt):
    spec, ip, ix, dv = small
    A = sp.csr_matrix((dv.astype(dt), ix, ip), shape=(3000, 900))
    mask = s.synth.make_mask(900, 300, seed=7) if masked else None
    n_eff = 300 if masked else 900
    om = s.synth.make_omega(n_eff, 30, seed=42, dtype=np.float64)
    R.set_threads()
    r = R.pca_fit(ip, ix, dv.astype(dt), 3000, 900, 20, om, mask=mask)
    ref = O.sparse_pca_fit(A.astype(np.float64), 20, omega=om, mask=mask, n_oversamples=10, n_power_iterations=7)
    tol_s, tol_a = (1e-10, 1e-6) if dt == np.float64 else (2e-5, 1e-3)
    assert O.rel_err(r.singular_values, ref.singular_values) < tol_s
    assert O.largest_principal_angle(r.components, ref.components) < tol_a
    assert np.allclose(r.mean, ref.mean, rtol=1e-12, atol=0)
    assert abs(r.total_var - ref.total_var) < 1e-10 * abs(ref.total_var)
    assert np.allclose(r.explained_variance, ref.explained_variance, rtol=10 * tol_s)
    # signs (svd_flip) and the projection of the fitted rows
    ex = O.transform(A, r.components, r.mean, mask=mask, mode=O.EXACT)
    assert np.abs(r.scores - ex).max() < (1e-9 if dt == np.float64 else 2e-4) * np.abs(ex).max()
    if dt == np.float64:
        assert np.abs(np.abs(r.components) - np.abs(ref.components)).max() < 1e-6
        j = np.argmax(np.abs(r.components), axis=1)
        assert (r.components[np.arange(20), j] > 0).all()


def test_cpu_ref_col_sums(small):
    spec, ip, ix, dv = small
    assert np.allclose(R.col_sums_f32(ip, ix, dv, 3000, 900), O.sum_col(ip, ix, dv, 900, np.float64), rtol=1e-12)
    assert np.allclose(R.col_sums_f32(ip, ix, dv, 3000, 900, squared=True), O.sum_col_squared(ip, ix, dv, 900, np.float64),
                       rtol=1e-12)


def test_cpu_ref_uncentred_and_rank_clamp():
    rng = np.random.default_rng(3)
    A = sp.random(400, 40, density=0.2, random_state=5, format="csr", dtype=np.float64)
    om = rng.standard_normal((40, 40))       # rank clamps to 40 columns... n_components 30 + 10 oversamples
    r = R.pca_fit(A.indptr, A.indices, A.data, 400, 40, 30, om, center=False)
    ref = O.sparse_pca_fit(A, 30, omega=om, center=False)
    assert O.rel_err(r.singular_values, ref.singular_values) < 1e-9
    assert abs(r.total_var - ref.total_var) < 1e-9 * ref.total_var


def test_cpu_ref_preprocess_matches_numpy_oracle(small):
    spec, ip, ix, dv = small
    v = dv.copy()
    ssum, ssq = R.preprocess_f32(ip, ix, v, 3000, 900, 1e4)
    rs = O.sum_row(ip, ix, dv, 3000)
    ref = O.log1p_normalize(O.normalize(ip, ix, dv, rs, 1e4, O.ROW))
    assert np.allclose(v, ref, rtol=3e-6, atol=0)
    assert np.allclose(ssum, O.sum_col(ip, ix, ref, 900, np.float64), rtol=1e-5)
    assert np.allclose(ssq, O.sum_col_squared(ip, ix, ref, 900, np.float64), rtol=1e-5)
