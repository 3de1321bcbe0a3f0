"""oracle/cholqr_ref.py — numpy restatement of the CholeskyQR variants the CUDA path uses in place of nalgebra's Householder
`qr()` inside single-svdlib's randomized_svd (normaliser `PowerIterationNormalizer::QR`; call sites
/root/reference/src/dimred/pca/sparse/mod.rs:170-180, src/dimred/pca/sparse_masked/mod.rs:341-351).  TEST INFRASTRUCTURE ONLY: nothing
under single-algebra_b200/ imports this file; tests/test_oracle_cholqr.py pins it against numpy's Householder QR / dense SVD and
tests/test_gpu_fused_side.py checks the CUDA kernels (csrc/dense.cu: chol_inv_block, zside_solve_kernel) on the same inputs.

Parity status: the reference holds no golden vectors for its normaliser (the arithmetic lives in the un-vendored single-svdlib /
nalgebra); what is pinned here is the mathematical contract — an orthonormal basis of the same column space, R upper triangular with
Y = Q R — and the behaviour on rank-deficient panels, where Householder QR completes the basis with arbitrary orthonormal vectors and
this scheme returns zero columns instead."""
import numpy as np

FLOOR_REL = 1e-13       # floor pass: pivot against the largest diagonal entry of the Gram matrix (dense.cu: floor_piv)
DROP_REL = 1e-5         # drop pass: pivot against the column's own squared norm (scale-free)


def chol_inv(G, drop=False):
    """Cholesky factor L (G = L L^T) and L^{-1} of a Gram matrix, column by column, with the dependent-column rules of
    dense.cu::chol_inv_block: a pivot that is not safely positive marks a column that depends on the ones before it; its
    sub-diagonal column of L is zero (no trailing update) and either its pivot is floored (drop=False) or the column is dropped
    (zero column of L, zero row of L^{-1}).  Returns (L, L^{-1}, number of dependent columns)."""
    G = np.asarray(G, dtype=np.float64)
    n = G.shape[0]
    A = G.copy()
    L = np.zeros_like(G)
    Li = np.zeros_like(G)
    diag0 = G.diagonal().copy()
    md = diag0.max() if n else 0.0
    floor = (md if md > 0 else 1.0) * FLOOR_REL
    n_dep = 0
    for j in range(n):
        p = A[j, j]
        dep = (not p > DROP_REL * diag0[j]) if drop else (not p > floor)
        if dep:
            n_dep += 1
            p, d = (0.0, 0.0) if drop else (floor, 1.0 / np.sqrt(floor))
        else:
            d = 1.0 / np.sqrt(p)
        L[j, j] = p * d
        if not dep:
            L[j + 1:, j] = A[j + 1:, j] * d
            A[j + 1:, j + 1:] -= np.outer(L[j + 1:, j], L[j + 1:, j])
        row = np.zeros(n)
        row[j] = 1.0
        row -= L[j, :j] @ Li[:j, :]
        Li[j, :] = row * d
    return L, Li, n_dep


def cholqr2(Y):
    """CholeskyQR2 as dense.cu::cholqr2 runs it (f64 Gram of the panel in its own precision, first pass floors, last pass
    drops): returns (Q in Y's dtype, R = R2 R1 in f64, dependent columns seen in the last pass)."""
    dt = Y.dtype
    Q = Y.copy()
    R = np.eye(Y.shape[1])
    n_dep = 0
    for ps in range(2):
        G = Q.astype(np.float64).T @ Q.astype(np.float64)
        L, Li, n_dep = chol_inv(G, drop=(ps == 1))
        Q = (Q @ Li.T.astype(dt)).astype(dt)
        R = L.T @ R
    return Q, R, n_dep


def small_side_two_step(Z0, Gy):
    """The fused half step of the power iteration (dense.cu::zside_solve_kernel, two-step mode): from the Gram of the tall panel
    Gy = Y^T Y and the raw product Z0 = A_c^T Y, the 64 x 64 matrix M with Z0 M = orth(Z0 R1^{-1}), formed WITHOUT touching the
    panel: R1 = chol(Gy), G1 = R1^{-T} (Z0^T Z0) R1^{-1}, R2 = chol(G1), M = R1^{-1} R2^{-1}."""
    GZ = Z0.astype(np.float64).T @ Z0.astype(np.float64)
    _, Li1, _ = chol_inv(Gy)
    Ri1 = Li1.T
    G1 = Ri1.T @ GZ @ Ri1
    _, Li2, _ = chol_inv(G1)
    return Ri1 @ Li2.T
