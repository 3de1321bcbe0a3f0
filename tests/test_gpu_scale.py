"""Full-size checks through size-independent properties (BASELINE configs 2 and 3): device-generated
count matrices, adjoint identity of the two sparse products, linearity, checksum of checksums,
orthonormal components, and a host spot check of the generator."""
import numpy as np
import pytest

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cfg2(salg, ctx):
    spec = salg.synth.make_spec(100_000, 20_000, density=0.07, seed=42)
    return spec, salg.synth_device(spec, dtype=np.float32, ctx=ctx)


def test_config2_generator_spot_check_and_checksums(salg, ctx, cfg2):
    spec, d = cfg2
    assert abs(d.nnz / (100_000 * 20_000) - 0.07) < 0.005
    shard = salg.synth_device(spec, 99_000, 64, dtype=np.float32, ctx=ctx)
    off, idx, val = shard.download()
    ip, ix, dv = salg.synth.generate_rows(spec, 99_000, 99_064, dtype=np.float32)
    assert np.array_equal(off.astype(np.int64), ip) and np.array_equal(idx.astype(np.int64), ix) and np.array_equal(val, dv)
    # checksum of checksums: integer counts => f64 totals are exact
    s, q, cnt, var = d.col_stats()
    rs = d.sum_row().astype(np.float64)
    assert s.sum() == rs.sum()
    assert cnt.sum() == d.nnz
    assert np.all(q >= s)            # counts >= 1  =>  x^2 >= x


def test_config2_products_adjoint_and_linear(salg, ctx, cfg2):
    spec, d = cfg2
    rng = np.random.default_rng(0)
    X = rng.standard_normal((20_000, 60)).astype(np.float32)
    Y = rng.standard_normal((100_000, 60)).astype(np.float32)
    mu = (d.sum_col() / 100_000).astype(np.float32)
    AX = salg.op_spmm(d, X, mu=mu).astype(np.float64)
    AtY = salg.op_spmm(d, Y, mu=mu, transposed=True).astype(np.float64)
    lhs = np.sum(AX * Y, dtype=np.float64)          # <A_c X, Y>
    rhs = np.sum(X * AtY, dtype=np.float64)         # <X, A_c^T Y>
    assert abs(lhs - rhs) < 1e-5 * np.sqrt(np.sum(AX ** 2) * np.sum(Y.astype(np.float64) ** 2))
    X2 = rng.standard_normal((20_000, 60)).astype(np.float32)
    A12 = salg.op_spmm(d, X + X2).astype(np.float64)
    A1 = salg.op_spmm(d, X).astype(np.float64)
    A2 = salg.op_spmm(d, X2).astype(np.float64)
    assert np.abs(A12 - A1 - A2).max() < 1e-4 * np.abs(A12).max()
    # centred product annihilates the all-ones direction: 1^T (A_c X) = 0
    assert np.abs(AX.sum(axis=0)).max() < 1e-3 * np.abs(AX).sum(axis=0).max()


def test_config2_fit_properties(salg, ctx, cfg2):
    spec, d = cfg2
    om = salg.synth.make_omega(20_000, 60, seed=42, dtype=np.float32)
    pca = salg.SparsePCABuilder().n_components(50).svd_method(
        salg.SVDMethod.Random(10, 7, salg.PowerIterationNormalizer.QR)).build()
    scores = pca.fit_transform(d, omega=om)
    V = pca.components_.astype(np.float64)
    assert np.abs(V @ V.T - np.eye(50)).max() < 1e-4
    s = pca.singular_values_
    assert np.all(np.diff(s) <= 1e-6 * s[0]) and np.all(s > 0)
    assert np.allclose(pca.explained_variance_, s ** 2 / (100_000 - 1), rtol=1e-5)
    # scores columns: centred, norms equal the singular values of the projected operator
    sc = scores.astype(np.float64)
    assert np.abs(sc.mean(axis=0)).max() < 1e-3 * np.abs(sc).max()
    assert np.allclose(np.linalg.norm(sc, axis=0), s, rtol=5e-3)
    assert pca.numeric_flags() == 0
    # idempotence: same inputs, same Omega => same singular values (atomics only reorder f32 partial rows)
    pca2 = salg.SparsePCABuilder().n_components(50).svd_method(
        salg.SVDMethod.Random(10, 7, salg.PowerIterationNormalizer.QR)).build()
    pca2.fit(d, omega=om)
    assert O.rel_err(pca2.singular_values_, s) < 1e-5


def test_config3_masked_fit_properties(salg, ctx):
    """BASELINE config 3 at full size: 1M x 30k at 7 %, 2000-gene mask, f32, randomized, k=50, q=7."""
    spec = salg.synth.make_spec(1_000_000, 30_000, density=0.07, seed=42)
    d = salg.synth_device(spec, dtype=np.float32, ctx=ctx)
    assert abs(d.nnz / 3e10 - 0.07) < 0.005
    mask = salg.synth.make_mask(30_000, 2_000, seed=7)
    comp = d.select_columns(mask)
    s_all, q_all, cnt, _ = d.col_stats()
    assert comp.nnz == int(cnt[mask].sum())           # compaction keeps exactly the kept columns' entries
    assert np.array_equal(comp.col_stats()[0], s_all[mask])
    comp.free()
    om = salg.synth.make_omega(2_000, 60, seed=42, dtype=np.float32)
    pca = salg.MaskedSparsePCABuilder().n_components(50).mask(mask.tolist()).svd_method(
        salg.SVDMethod.Random(10, 7, salg.PowerIterationNormalizer.QR)).build()
    pca.fit(d, omega=om)
    V = pca.components_.astype(np.float64)
    assert V.shape == (50, 2000) and pca.mean_.shape == (30_000,)
    assert np.abs(V @ V.T - np.eye(50)).max() < 1e-4
    assert np.all(np.diff(pca.singular_values_) <= 0)
    assert np.allclose(pca.mean_, s_all / 1e6, rtol=1e-5)
    # parity on a host-regenerable sub-sample is covered by bench.py's cpu_baseline leg and the small tests
    d.free()
