"""Host-side logic that needs no GPU: builders, row partitioning, generator determinism."""
import numpy as np
import scipy.sparse as sp

import single_algebra_b200 as s
from conftest import planted_counts


def test_builder_defaults_and_fluent_setters():
    # pca/sparse/mod.rs:388-403, pca/sparse_masked/mod.rs:51-67
    b = s.SparsePCABuilder.new()
    p = b.build()
    assert (p.n_components, p.alpha, p.tolerance, p.random_seed, p.center, p.verbose) == (50, 1.0, 1e-6, 42, True, False)
    assert p.svdmethod == s.SVDMethod.default() == s.SVDMethod.Lanczos
    m = s.SVDMethod.Random(10, 7, s.PowerIterationNormalizer.LU)
    p = s.SparsePCABuilder().n_components(7).alpha(0.5).tolerance(1e-3).random_seed(9).center(False).verbose(True) \
        .svd_method(m).build()
    assert (p.n_components, p.alpha, p.tolerance, p.random_seed, p.center, p.verbose, p.svdmethod) == \
        (7, 0.5, 1e-3, 9, False, True, m)
    mp = s.MaskedSparsePCABuilder().mask([True, False, True]).n_components(2).build()
    assert mp.mask.tolist() == [True, False, True] and mp._masked
    # constructor form: SparsePCA::new(n_components, alpha, tollerance, random_seed, center, verbose, svdmethod)
    q = s.SparsePCA(5, 1.0, None, None, True, False, m)
    assert q.tolerance == 1e-6 and q.random_seed == 42
    prm = q._params(keep_scores=True)
    assert (prm.n_components, prm.svd_method, prm.n_oversamples, prm.n_power_iterations, prm.normalizer,
            prm.center, prm.random_seed, prm.keep_scores) == (5, 1, 10, 7, 1, 1, 42, 1)


def test_not_fitted_errors():
    import pytest
    p = s.SparsePCABuilder().build()
    for fn in (p.feature_importances, p.explained_variance_ratio, p.cumulative_explained_variance_ratio):
        with pytest.raises(s.SalgError) as e:
            fn()
        assert str(e.value) == "Model must be fitted first!"      # pca/sparse/mod.rs:299, 316


def test_masked_transform_checks_the_mask_length_before_the_fitted_state():
    """pca/sparse_masked/mod.rs:440-444 comes before :449."""
    import pytest
    mp = s.MaskedSparsePCABuilder().mask([True, False, True]).build()
    x = s.CsrMatrix(2, 4, np.array([0, 1, 2], np.uint64), np.array([0, 3], np.uint64), np.array([1.0, 2.0]))
    with pytest.raises(s.SalgError) as e:
        mp.transform(x)
    assert e.value.code == 2 and "mask vector length" in str(e.value)


def test_partition_rows_by_nnz_balances_entries():
    A = planted_counts(1000, 50, seed=1)
    A = sp.vstack([A, sp.csr_matrix((200, 50))]).tocsr()      # trailing empty rows
    for n in (1, 2, 3, 8):
        parts = s.dist.partition_rows_by_nnz(A.indptr, n)
        assert parts[0][0] == 0 and parts[-1][1] == A.shape[0]
        assert all(parts[i][1] == parts[i + 1][0] for i in range(n - 1))
        nnz = [int(A.indptr[b] - A.indptr[a]) for a, b in parts]
        assert sum(nnz) == A.nnz
        assert max(nnz) - min(nnz) <= 2 * int(np.diff(A.indptr).max()) + 1
    assert s.dist.partition_rows_by_nnz(np.zeros(11, np.int64), 4) == [(0, 2), (2, 5), (5, 7), (7, 10)]
    assert s.dist.partition_rows_even(10, 4) == [(0, 2), (2, 5), (5, 7), (7, 10)]


def test_shard_csr_rebases_offsets():
    A = planted_counts(100, 20, seed=2)
    off, idx, val = s.dist.shard_csr(A.indptr, A.indices, A.data, 30, 70)
    B = sp.csr_matrix((val, idx, off), shape=(40, 20))
    assert (B != A[30:70]).nnz == 0


def test_generator_is_counter_based_and_hits_density():
    spec = s.synth.make_spec(3000, 800, density=0.07, seed=42)
    ip, ix, dv = s.synth.generate_rows(spec, 0, 600)
    ip2, ix2, dv2 = s.synth.generate_rows(spec, 300, 600)      # any row range regenerates identically
    a, b = ip[300], ip[600]
    assert np.array_equal(ix[a:b], ix2) and np.array_equal(dv[a:b], dv2)
    assert np.array_equal(ip[300:] - a, ip2)
    dens = ip[-1] / (600 * 800)
    assert abs(dens - 0.07) < 0.01
    assert dv.min() >= 1 and dv.dtype == np.float32
    assert np.all(np.diff(ix[ip[5]:ip[6]]) > 0)
    om = s.synth.make_omega(10, 4)
    assert np.array_equal(om, np.random.Generator(np.random.PCG64(42)).standard_normal((10, 4)).astype(np.float32))
    mk = s.synth.make_mask(30000, 2000)
    assert mk.sum() == 2000 and mk.dtype == bool
