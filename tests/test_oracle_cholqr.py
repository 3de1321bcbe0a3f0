"""CPU checks of oracle/cholqr_ref.py (the numpy restatement of the CholeskyQR variants of csrc/dense.cu) against numpy's
Householder QR and dense SVD: full-rank panels, ill-conditioned f32 panels, exactly rank-deficient panels (the regression input of
tests/test_gpu_fused_side.py) and the algebra of the fused small-side step."""
import numpy as np

from oracle import cholqr_ref as C


def test_cholqr2_matches_householder_on_full_rank_panels():
    rng = np.random.default_rng(0)
    Y = rng.standard_normal((5000, 60)) @ np.diag(np.logspace(0, 3, 60)) @ rng.standard_normal((60, 60))
    for dt, tol in ((np.float64, 1e-12), (np.float32, 5e-5)):
        Q, R, n_dep = C.cholqr2(Y.astype(dt))
        Q = Q.astype(np.float64)
        assert n_dep == 0
        assert np.abs(Q.T @ Q - np.eye(60)).max() < tol
        assert np.abs(Q @ R - Y).max() < tol * np.abs(Y).max() * 10
        assert np.allclose(R, np.triu(R))
        qh, _ = np.linalg.qr(Y)
        assert np.abs(qh @ (qh.T @ Q) - Q).max() < tol * 100        # same column space as Householder QR


def test_rank_deficient_panel_is_dropped_not_amplified():
    """3000 rows that are copies of 12 profiles, centred: rank 11, sketch of 30 columns.  The flooring rule of rounds 1-2 (keep the
    sub-diagonal column of a floored pivot) overflows on this input; the current rule drops the dependent columns and the
    singular values of Q^T A are those of A."""
    rng = np.random.default_rng(3)
    base = (rng.random((12, 400)) < 0.2) * rng.integers(1, 6, size=(12, 400))
    D = base[rng.integers(0, 12, size=3000)].astype(np.float64)
    Dc = D - D.mean(axis=0)
    s_true = np.linalg.svd(Dc, compute_uv=False)
    om = np.random.default_rng(42).standard_normal((400, 30))
    for dt, tol in ((np.float64, 1e-10), (np.float32, 1e-5)):
        Y = (Dc @ om).astype(dt)
        Q, _, n_dep = C.cholqr2(Y)
        assert n_dep in (18, 19) and np.isfinite(Q).all()     # (a noise column outside the range may survive, orthonormal)
        norms = np.linalg.norm(Q.astype(np.float64), axis=0)
        assert np.sum(norms > 0.5) == 30 - n_dep and np.all((norms < 1e-12) | (np.abs(norms - 1) < 1e-3))
        Z = (Dc.T @ Q.astype(np.float64)).astype(dt)
        _, Rz, _ = C.cholqr2(Z)
        s = np.linalg.svd(Rz, compute_uv=False)
        assert np.max(np.abs(s[:11] - s_true[:11]) / s_true[:11]) < tol
        assert np.all(s[11:] < 1e-6 * s[0])
    # the old rule: floored pivots whose sub-diagonal columns keep updating the trailing matrix grow geometrically
    G = (Dc @ om).astype(np.float32).astype(np.float64)
    G = G.T @ G
    A = G.copy()
    fl = G.diagonal().max() * C.FLOOR_REL
    with np.errstate(all="ignore"):
        for j in range(30):
            p = A[j, j] if A[j, j] > fl else fl
            col = A[j + 1:, j] / np.sqrt(p)
            A[j + 1:, j + 1:] -= np.outer(col, col)
    assert not np.isfinite(A).all()


def test_small_side_two_step_is_the_explicit_chain():
    """M = R1^{-1} R2^{-1} from the two Gram matrices alone reproduces orth(Z0 R1^{-1}) of the explicit chain."""
    rng = np.random.default_rng(1)
    A = rng.poisson(0.3, size=(4000, 500)).astype(np.float64)
    A -= A.mean(axis=0)
    Z2_prev, _ = np.linalg.qr(rng.standard_normal((500, 60)))
    Y = (A @ Z2_prev).astype(np.float32)
    Z0 = (A.T @ Y.astype(np.float64)).astype(np.float32)
    Gy = Y.astype(np.float64).T @ Y.astype(np.float64)
    M = C.small_side_two_step(Z0, Gy)
    Z2 = Z0.astype(np.float64) @ M
    assert np.abs(Z2.T @ Z2 - np.eye(60)).max() < 1e-9
    # explicit chain: Z1 = Z0 R1^{-1}, Z2 = Z1 R2^{-1} with R2 from the Gram of Z1 itself
    _, Li1, _ = C.chol_inv(Gy)
    Z1 = Z0.astype(np.float64) @ Li1.T
    _, Li2, _ = C.chol_inv(Z1.T @ Z1)
    assert np.abs(Z2 - Z1 @ Li2.T).max() < 1e-9
    qh, _ = np.linalg.qr(Z0.astype(np.float64))
    assert np.abs(qh @ (qh.T @ Z2) - Z2).max() < 1e-9               # same column space as Householder QR of Z0
