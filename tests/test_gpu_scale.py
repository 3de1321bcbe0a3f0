"""Full-size checks through size-independent properties (BASELINE configs 2, 3, 4 and one GPU's share of 5): device-generated
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


def test_config4_lanczos_f64_properties(salg, ctx):
    """BASELINE config 4 at full size: SparsePCA f64, SVDMethod::Lanczos, 250k x 20k at 7 %, 50 components (uncentred
    operator).  Properties: orthonormal right vectors, descending sigma, ||A v_i|| = sigma_i through the
    independent product kernel, and agreement of the leading sigma with a randomized fit of the uncentred operator."""
    spec = salg.synth.make_spec(250_000, 20_000, density=0.07, seed=42)
    d = salg.synth_device(spec, dtype=np.float64, ctx=ctx)
    pca = salg.SparsePCABuilder().n_components(50).center(False).build()          # default svd_method = Lanczos
    pca.fit(d)
    V = pca.components_
    s = pca.singular_values_
    assert V.shape == (50, 20_000) and V.dtype == np.float64
    assert np.abs(V @ V.T - np.eye(50)).max() < 1e-10
    assert np.all(np.diff(s) <= 0) and np.all(s > 0)
    Vp = np.zeros((20_000, 60))
    Vp[:, :50] = V.T
    AV = salg.op_spmm(d, Vp)[:, :50]
    assert np.allclose(np.linalg.norm(AV, axis=0), s, rtol=1e-9)
    assert np.abs(AV.T @ AV - np.diag(s ** 2)).max() < 1e-8 * s[0] ** 2              # left vectors orthogonal too
    om = salg.synth.make_omega(20_000, 60, seed=42, dtype=np.float64)
    rnd = salg.SparsePCABuilder().n_components(50).center(False).svd_method(
        salg.SVDMethod.Random(10, 7, salg.PowerIterationNormalizer.QR)).build()
    rnd.fit(d, omega=om)
    assert O.rel_err(rnd.singular_values_[:10], s[:10]) < 1e-6
    d.free()


def test_config5_shard_preprocessing_properties(salg, ctx):
    """One GPU's share of BASELINE config 5 (500k x 33k f32 at 7 %): normalize(ROW, 1e4) + log1p + column statistics.
    Properties: row sums hit the target, expm1 undoes log1p on a host-regenerated sample, the fused pass equals the chain,
    checksum of checksums."""
    spec = salg.synth.make_spec(4_000_000, 33_000, density=0.07, seed=42)
    r0, n = 1_500_000, 500_000
    d = salg.synth_device(spec, r0, n, dtype=np.float32, ctx=ctx)
    rs = d.sum_row()
    assert rs.shape == (n,) and np.all(rs >= 0)
    d.normalize(rs, 1e4, salg.Direction.ROW)
    rs2 = d.sum_row().astype(np.float64)
    nz = rs > 0
    assert np.abs(rs2[nz] - 1e4).max() < 1e4 * 2e-5 and np.all(rs2[~nz] == 0)
    d.log1p_normalize()
    s_chain, q_chain = d.sum_col_and_squared()
    # host-regenerated sample rows through the same chain in f64
    ip, ix, dv = salg.synth.generate_rows(spec, r0 + 1234, r0 + 1234 + 48, dtype=np.float32)
    sub = salg.synth_device(spec, r0 + 1234, 48, dtype=np.float32, ctx=ctx)
    sub.normalize(sub.sum_row(), 1e4, salg.Direction.ROW)
    sub.log1p_normalize()
    _, _, got = sub.download()
    ref = np.empty(len(dv))
    for r in range(48):
        a, b = ip[r], ip[r + 1]
        tot = dv[a:b].astype(np.float64).sum()
        ref[a:b] = np.log1p(dv[a:b].astype(np.float64) * (1e4 / tot)) if tot > 0 else dv[a:b]
    assert np.abs(got - ref).max() < 2e-6 * np.abs(ref).max()
    # fused preprocess on a fresh copy of the shard equals the chain
    d2 = salg.synth_device(spec, r0, n, dtype=np.float32, ctx=ctx)
    s_f, q_f = d2.preprocess(1e4)
    assert np.allclose(s_f, s_chain, rtol=2e-5, atol=1e-3) and np.allclose(q_f, q_chain, rtol=2e-5, atol=1e-3)
    assert abs(float(np.sum(s_f, dtype=np.float64)) - float(np.sum(d2.sum_row(), dtype=np.float64))) < 1e-5 * float(np.sum(s_f, dtype=np.float64))
    d.free(); d2.free(); sub.free()
