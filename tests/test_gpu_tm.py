"""TMEM-operand products (csrc/tm.cu, spmm_impl 'tm'): the sparse operand is expanded from the quad-mask tile format into
tensor memory and contracted by tcgen05.mma with A from TMEM.  Same checks as the tile-densified generation: products
against an f64 reference (2e-5 of the largest entry), ragged / dense / tiny shapes, many work items per CTA, and whole fits
against the oracle."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import planted_counts
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _ref_products(A, X, mu, transposed):
    A = A.astype(np.float64)
    X = X.astype(np.float64)
    if not transposed:
        Y = A @ X
        return Y - (mu.astype(np.float64) @ X)[None, :] if mu is not None else Y
    Z = A.T @ X
    return Z - mu.astype(np.float64)[:, None] * X.sum(axis=0)[None, :] if mu is not None else Z


@pytest.fixture()
def tm_ctx(ctx):
    ctx.set_spmm_impl("tm")      # (the default; explicit so the file does not depend on SALG_SPMM_IMPL)
    yield ctx


def _check(salg, ctx, A, k=60, seed=0, tol=2e-5):
    rng = np.random.default_rng(seed)
    d = salg.CsrMatrix.from_scipy(A, ctx).to_device()
    mu = np.asarray(A.mean(axis=0)).ravel().astype(np.float32)
    for transposed in (False, True):
        X = rng.standard_normal((A.shape[0] if transposed else A.shape[1], k)).astype(np.float32)
        for m in (None, mu):
            got = salg.op_spmm(d, X, mu=m, transposed=transposed)
            ref = _ref_products(A, X, m, transposed)
            scale = np.abs(_ref_products(A, X, None, transposed)).max()
            err = np.abs(got - ref).max()
            assert err <= tol * scale, (A.shape, transposed, m is None, err / scale)


def test_tm_products_counts_and_general(salg, tm_ctx):
    """Exact-fp16 operator (raw counts: one term) and a general float operator (two terms); shapes that are not multiples
    of the 128 x 128 tile; empty rows / columns, one long row."""
    rng = np.random.default_rng(5)
    for general in (False, True):
        A = planted_counts(1000 + 37, 300 + 11, seed=31, dtype=np.float32)
        D = A.toarray()
        D[5] = 0
        D[17] = np.arange(D.shape[1]) % 7 + 1
        D[:, 3] = 0
        D[-1] = 0
        A = sp.csr_matrix(D.astype(np.float32))
        if general:
            A.data = (A.data * (1 + rng.random(A.nnz))).astype(np.float32)
        _check(salg, tm_ctx, A, seed=int(general))
        for k in (1, 64):
            _check(salg, tm_ctx, A, k=k, seed=k)


def test_tm_products_dense_and_tiny(salg, tm_ctx):
    """Fully dense 128 x 128 tiles hold 4096 quads, more than a ring slot: the expanders then read the quads from global
    memory.  Tiny shapes exercise the padding (rows < 128, columns < 128)."""
    rng = np.random.default_rng(9)
    for shape, dens in (((300, 200), 1.0), ((257, 130), 0.6), ((5, 3), 1.0), ((1, 70), 0.5), ((129, 1), 1.0)):
        D = rng.integers(1, 9, size=shape).astype(np.float32) * (rng.random(shape) < dens)
        D[0, 0] = 3.0
        _check(salg, tm_ctx, sp.csr_matrix(D), seed=shape[0])


def test_tm_products_many_groups_per_cta(salg, tm_ctx):
    """Tall: more row-block pairs than SMs (accumulator sets alternate, ring and TMEM buffers wrap many times).
    Wide: many column groups x row ranges in A^T Y (several work items per CTA, flushes between them)."""
    A = planted_counts(60_000, 200, seed=3, dtype=np.float32)
    _check(salg, tm_ctx, A, seed=1)
    A = planted_counts(2500, 6000, seed=4, density=0.03, dtype=np.float32)
    _check(salg, tm_ctx, A, seed=2)


def test_tm_aty_three_items_per_cta_with_single_block_group(salg, tm_ctx):
    """12 000 x 20 000: 157 column blocks = 78 column groups of two + one group of ONE block, 5 row ranges: 395 work items,
    three per CTA.  The MMA issuer without passes in the single-block group must still wait for the epilogue of the group
    before (regression: it ran a group ahead and the parity of its next wait aliased -> deadlock)."""
    spec = salg.synth.make_spec(12_000, 20_000, density=0.07, seed=42)
    d = salg.synth_device(spec, dtype=np.float32, ctx=tm_ctx)
    rng = np.random.default_rng(3)
    for transposed in (False, True):
        X = rng.standard_normal((12_000 if transposed else 20_000, 60)).astype(np.float32)
        tm_ctx.set_spmm_impl("chunk")
        ref = salg.op_spmm(d, X, transposed=transposed)
        tm_ctx.set_spmm_impl("tm")
        got = salg.op_spmm(d, X, transposed=transposed)
        assert np.abs(got - ref).max() <= 4e-5 * np.abs(ref).max(), transposed
    d.free()


def test_tm_matches_dense_tile_generation(salg, ctx):
    """Both tensor-core generations compute exact fp16 x fp16 products with f32 accumulation: they agree to f32 rounding."""
    A = planted_counts(3000, 900, seed=8, dtype=np.float32)
    rng = np.random.default_rng(0)
    d = salg.CsrMatrix.from_scipy(A, ctx).to_device()
    for transposed in (False, True):
        X = rng.standard_normal((A.shape[0] if transposed else A.shape[1], 60)).astype(np.float32)
        ctx.set_spmm_impl("tc")
        a = salg.op_spmm(d, X, transposed=transposed)
        ctx.set_spmm_impl("tm")
        b = salg.op_spmm(d, X, transposed=transposed)
        assert np.abs(a - b).max() <= 2e-6 * np.abs(a).max()


@pytest.mark.parametrize("masked", [False, True])
def test_tm_fit_against_oracle(salg, tm_ctx, masked):
    A = planted_counts(6000, 1200, seed=21, dtype=np.float32)
    mask = salg.synth.make_mask(1200, 400, seed=7) if masked else None
    n_eff = 400 if masked else 1200
    om = salg.synth.make_omega(n_eff, 40, seed=42, dtype=np.float32)
    ref = O.sparse_pca_fit(A.astype(np.float64), 30, omega=om.astype(np.float64), mask=mask, n_oversamples=10,
                           n_power_iterations=7)
    rnd = salg.SVDMethod.Random(10, 7, salg.PowerIterationNormalizer.QR)
    if masked:
        pca = salg.MaskedSparsePCABuilder().n_components(30).mask(mask.tolist()).svd_method(rnd).build()
    else:
        pca = salg.SparsePCABuilder().n_components(30).svd_method(rnd).build()
    x = salg.CsrMatrix.from_scipy(A, tm_ctx)
    scores = pca.fit_transform(x, omega=om)
    assert O.rel_err(pca.singular_values_, ref.singular_values) < 1e-4
    assert O.largest_principal_angle(pca.components_, ref.components) < 1e-3
    ex = O.transform(A, pca.components_, pca.mean_, center=True, mask=mask, mode=O.EXACT)
    assert np.abs(scores - ex).max() < 2e-4 * np.abs(ex).max()
