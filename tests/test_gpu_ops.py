"""Operator-level parity: centred sparse x panel products, CholeskyQR2, the one-CTA Jacobi SVD."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import planted_counts

pytestmark = pytest.mark.gpu


def _ref_products(A, X, mu, transposed):
    A = A.astype(np.float64)
    X = X.astype(np.float64)
    if not transposed:
        Y = A @ X
        return Y - (mu.astype(np.float64) @ X)[None, :] if mu is not None else Y
    Z = A.T @ X
    return Z - mu.astype(np.float64)[:, None] * X.sum(axis=0)[None, :] if mu is not None else Z


@pytest.mark.parametrize("dtype,tol", [(np.float32, 2e-5), (np.float64, 1e-12)])
@pytest.mark.parametrize("k", [1, 60, 64])
def test_spmm_matches_oracle(salg, ctx, dtype, tol, k):
    A = planted_counts(3000, 500, seed=11, dtype=dtype)
    # ragged structure: empty rows, one very long row (spans several work chunks), empty columns
    D = A.toarray()
    D[5] = 0
    D[17] = np.arange(500) % 7 + 1
    D[:, 3] = 0
    D[-1] = 0
    A = sp.csr_matrix(D.astype(dtype))
    d = salg.CsrMatrix.from_scipy(A, ctx).to_device()
    rng = np.random.default_rng(k)
    mu = np.asarray(A.mean(axis=0)).ravel().astype(dtype)
    for transposed in (False, True):
        X = rng.standard_normal((A.shape[0] if transposed else A.shape[1], k)).astype(dtype)
        for m in (None, mu):
            got = salg.op_spmm(d, X, mu=m, transposed=transposed)
            ref = _ref_products(A, X, m, transposed)
            scale = np.abs(ref).max()
            assert np.abs(got - ref).max() <= tol * scale, (transposed, m is None)


def test_spmm_single_row_and_tiny(salg, ctx):
    A = sp.csr_matrix(np.array([[0.0, 2.0, 0.0, 1.0]]))
    d = salg.CsrMatrix.from_scipy(A, ctx).to_device()
    X = np.arange(8, dtype=np.float64).reshape(4, 2)
    assert np.allclose(salg.op_spmm(d, X), A @ X)
    Y = np.array([[3.0, -1.0]])
    assert np.allclose(salg.op_spmm(d, Y, transposed=True), A.T @ Y)


@pytest.mark.parametrize("dtype,tol", [(np.float32, 5e-6), (np.float64, 1e-13)])
def test_cholqr2(salg, ctx, dtype, tol):
    rng = np.random.default_rng(0)
    Y = (rng.standard_normal((5000, 60)) @ np.diag(np.logspace(0, 3, 60)) @ rng.standard_normal((60, 60))).astype(dtype)
    q, r = salg.op_cholqr2(Y, ctx)
    assert np.abs(q.astype(np.float64).T @ q.astype(np.float64) - np.eye(60)).max() < tol * 10
    assert np.abs(q.astype(np.float64) @ r - Y).max() < tol * np.abs(Y).max() * 10
    assert np.allclose(r, np.triu(r))
    # same column space as Householder QR
    qh, _ = np.linalg.qr(Y.astype(np.float64))
    assert np.abs(qh @ (qh.T @ q) - q).max() < tol * 100


@pytest.mark.parametrize("k", [1, 7, 60, 64])
def test_small_svd_matches_lapack(salg, ctx, k):
    rng = np.random.default_rng(k)
    a = rng.standard_normal((k, k)) @ np.diag(np.logspace(0, -6, k)) @ rng.standard_normal((k, k))
    u, s, vt = salg.op_small_svd(a, ctx)
    sr = np.linalg.svd(a, compute_uv=False)
    assert np.max(np.abs(s - sr) / sr[0]) < 1e-13
    assert np.all(np.diff(s) <= 0)
    assert np.abs(u @ np.diag(s) @ vt - a).max() < 1e-12 * np.abs(a).max()
    assert np.abs(u.T @ u - np.eye(k)).max() < 1e-10 and np.abs(vt @ vt.T - np.eye(k)).max() < 1e-10


def test_tc_products_match_chunk_kernels_and_f64(salg, ctx):
    """tcgen05 products vs the CUDA-core kernels vs f64: exact-fp16 operator (raw counts, one operator term) and a general
    float operator (two fp16 terms; the dense panel is always two fp16 terms = 22 significant bits, tc.cu:15-18), ragged
    shapes, both products."""
    rng = np.random.default_rng(5)
    for make_general in (False, True):
        A = planted_counts(1000 + 37, 300 + 11, seed=31, dtype=np.float32)      # not multiples of 128 / 64
        if make_general:
            A.data = (A.data * (1 + rng.random(A.nnz))).astype(np.float32)
        d = salg.CsrMatrix.from_scipy(A, ctx).to_device()
        mu = np.asarray(A.mean(axis=0)).ravel().astype(np.float32)
        for transposed in (False, True):
            X = rng.standard_normal((A.shape[0] if transposed else A.shape[1], 60)).astype(np.float32)
            ref = _ref_products(A, X, mu, transposed)
            ctx.set_spmm_impl("tc")
            got_tc = salg.op_spmm(d, X, mu=mu, transposed=transposed)
            ctx.set_spmm_impl("chunk")
            got_ch = salg.op_spmm(d, X, mu=mu, transposed=transposed)
            ctx.set_spmm_impl("tm")        # the context's default
            scale = np.abs(ref).max()
            assert np.abs(got_tc - ref).max() <= 2e-5 * scale, (make_general, transposed)
            assert np.abs(got_ch - ref).max() <= 2e-5 * scale
            # the two-term fp16 tensor-core product (exact products, f32 accumulation) is as accurate as the f32 FMA chain
            assert np.abs(got_tc - ref).max() <= 4 * np.abs(got_ch - ref).max() + 1e-6 * scale


def test_tc_products_dense_tiles_and_tiny_shapes(salg, ctx):
    """Tiles denser than a ring slot (fully dense blocks: 8192 entries per 128 x 64 tile) take the overflow path of
    the scatter role; tiny shapes exercise padding (rows < 128, columns < 64)."""
    rng = np.random.default_rng(9)
    for shape, dens in (((300, 200), 1.0), ((257, 130), 0.6), ((5, 3), 1.0), ((1, 70), 0.5)):
        D = rng.integers(1, 9, size=shape).astype(np.float32) * (rng.random(shape) < dens)
        D[0, 0] = 3.0
        A = sp.csr_matrix(D)
        d = salg.CsrMatrix.from_scipy(A, ctx).to_device()
        mu = np.asarray(A.mean(axis=0)).ravel().astype(np.float32)
        ctx.set_spmm_impl("tc")
        for transposed in (False, True):
            X = rng.standard_normal((shape[0] if transposed else shape[1], 60)).astype(np.float32)
            ref = _ref_products(A, X, mu, transposed)
            got = salg.op_spmm(d, X, mu=mu, transposed=transposed)
            # scale of the un-centred product (a one-row matrix is annihilated by the centring)
            scale = np.abs(_ref_products(A, X, None, transposed)).max()
            assert np.abs(got - ref).max() <= 2e-5 * scale, (shape, dens, transposed)
    ctx.set_spmm_impl("tm")                # the context's default


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1, 5), (127, 60), (128, 64), (1000, 33), (70_001, 60)])
def test_tall_gram_fused_pass(salg, ctx, shape):
    """tc_gram_prep_kernel: Gram + column sums of a tall f32 panel (fp16 two-term products on the tensor core, f32
    accumulation drained to f64 every 1024 rows) against an f64 reference; columns of very different scale."""
    rng = np.random.default_rng(5)
    Y = rng.standard_normal(shape).astype(np.float32) * np.logspace(0, -3, shape[1]).astype(np.float32)[None, :]
    Y[:, 0] += 0.5        # non-zero column mean
    g, cs, _ = salg.op_tall_gram(Y, ctx)
    Y64 = Y.astype(np.float64)
    ref = Y64.T @ Y64
    nrm = np.sqrt(np.outer(np.diag(ref), np.diag(ref)))
    assert np.abs(g - ref).max() <= 3e-6 * nrm.max()
    assert (np.abs(g - ref) <= 2e-5 * nrm + 1e-30).all()          # every entry relative to its own columns' norms
    assert np.allclose(cs, Y64.sum(axis=0), rtol=1e-6, atol=1e-6 * np.abs(Y64).sum(axis=0).max())
