"""CPU oracle for the sparse-PCA hot path of SingleRust/single-algebra 0.9.2.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and there only as the checker or as the
timed CPU baseline.  The product path (``single-algebra_b200``) never imports it.

Parity status
-------------
* Column/row sums, normalize, log1p: pinned against the reference's own known-answer
  tests (KAT-S1 ``src/sparse/csc.rs:1124-1152``, KAT-N1 ``src/sparse/csr.rs:1516-1550``,
  KAT-N2 ``src/sparse/csc.rs:1257-1301``, KAT-L1 ``src/sparse/csc.rs:1304-1314``) in
  ``tests/test_oracle_kat.py``.
* Randomized / Lanczos SVD, mask compaction: **parity unpinned**.  The arithmetic lives
  in the un-vendored dependency ``single-svdlib = 1.0.9`` (``Cargo.toml:37``,
  ``Cargo.lock:1393-1409``); the reference holds no golden vectors for it (its only PCA
  test asserts ``is_ok()``, ``src/dimred/pca/sparse/mod.rs:540-562``) and there is no
  Rust toolchain in this image, so the reference cannot be run.  These functions restate
  the published algorithm (Halko/Martinsson/Tropp Alg. 4.4 + 5.1 as used by
  scikit-learn's ``randomized_svd``, which the reference README credits) around the
  reference's own call sites, and are cross-checked against scikit-learn and a dense
  LAPACK SVD in ``tests/test_oracle_svd.py``.

Every function cites the reference lines it follows.
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sp

ROW = 0      # single_utilities::types::Direction::ROW
COLUMN = 1   # single_utilities::types::Direction::COLUMN


# --------------------------------------------------------------------------------------
# MatrixSum  (src/sparse/mod.rs:67-102; CSR impl src/sparse/csr.rs:259-312, 314-392, 558-608)
# --------------------------------------------------------------------------------------
def sum_col(indptr, indices, data, ncols, out_dtype=None):
    """``MatrixSum::sum_col`` for CSR (src/sparse/csr.rs:259-312): out[c] = sum of stored
    values whose column index is c.  The reference accumulates in the output type in a
    scheduler-dependent order (rayon chunks of 8192, :286-309); the oracle accumulates in
    f64 and casts, which is the order-independent value the tolerance is stated against."""
    out_dtype = out_dtype or data.dtype
    if len(data) == 0 or ncols == 0:           # csr.rs:269-271
        return np.zeros(ncols, dtype=out_dtype)
    acc = np.bincount(np.asarray(indices, dtype=np.int64),
                      weights=np.asarray(data, dtype=np.float64), minlength=ncols)
    return acc.astype(out_dtype)


def sum_col_squared(indptr, indices, data, ncols, out_dtype=None):
    """``MatrixSum::sum_col_squared`` (src/sparse/csr.rs:558-608): out[c] = sum of v*v."""
    out_dtype = out_dtype or data.dtype
    if len(data) == 0 or ncols == 0:
        return np.zeros(ncols, dtype=out_dtype)
    d = np.asarray(data, dtype=np.float64)
    acc = np.bincount(np.asarray(indices, dtype=np.int64), weights=d * d, minlength=ncols)
    return acc.astype(out_dtype)


def sum_row(indptr, indices, data, nrows, out_dtype=None):
    """``MatrixSum::sum_row`` (src/sparse/csr.rs:314-392): out[r] = sum of the row's stored
    values (the reference's 4-way / 8-chunk partial sums only change rounding order)."""
    out_dtype = out_dtype or data.dtype
    if nrows == 0:                              # csr.rs:323-325
        return np.zeros(0, dtype=out_dtype)
    d = np.asarray(data, dtype=np.float64)
    ip = np.asarray(indptr, dtype=np.int64)
    # per-row sums via reduceat over the non-empty rows (no global-cumsum cancellation)
    out = np.zeros(nrows, dtype=np.float64)
    nonempty = ip[1:] > ip[:-1]
    if len(d):
        starts = ip[:-1][nonempty]
        red = np.add.reduceat(d, starts) if len(starts) else np.zeros(0)
        # reduceat sums up to the next start; rows are contiguous so that is the row end
        out[nonempty] = red
    return out.astype(out_dtype)


# --------------------------------------------------------------------------------------
# Normalize / Log1P  (src/utils/mod.rs:6-17; CSR impls src/sparse/csr.rs:1013-1079)
# --------------------------------------------------------------------------------------
def normalize(indptr, indices, data, sums, target, direction):
    """``Normalize::normalize`` for CSR (src/sparse/csr.rs:1013-1068).

    scale[i] = target / sums[i] if sums[i] > 0 else 0           (:1021-1030)
    value is replaced by T(U(value) * scale[idx]) only where scale > 0  (:1041-1043, :1055-1061)
    COLUMN indexes the scale by column id, ROW by row id.  Arithmetic happens in the type
    of ``sums`` (U) and the product is cast back to the value type (T)."""
    sums = np.asarray(sums)
    U = sums.dtype
    T = data.dtype
    with np.errstate(divide="ignore", invalid="ignore"):
        scale = np.where(sums > 0, U.type(target) / sums, U.type(0)).astype(U)
    out = data.copy()
    if direction == COLUMN:
        s = scale[np.asarray(indices, dtype=np.int64)]
    elif direction == ROW:
        ip = np.asarray(indptr, dtype=np.int64)
        s = np.repeat(scale[: len(ip) - 1], np.diff(ip))
    else:
        raise ValueError("direction")
    m = s > 0
    out[m] = (data[m].astype(U) * s[m]).astype(T)
    return out


def log1p_normalize(data):
    """``Log1P::log1p_normalize`` (src/sparse/csr.rs:1070-1079): v <- ln(fl(1 + v)), two
    roundings in the value type — *not* ``ln_1p``."""
    one = data.dtype.type(1)
    return np.log((one + data).astype(data.dtype)).astype(data.dtype)


# --------------------------------------------------------------------------------------
# MaskedCSRMatrix (single-svdlib lanczos::masked; call site pca/sparse_masked/mod.rs:313)
# --------------------------------------------------------------------------------------
def mask_compact(indptr, indices, data, mask):
    """Column-subset view: keeps entries whose column has mask==True; compact id of an
    original column c is the number of True entries in mask[0..c].  Row structure and
    within-row order are preserved (equals scipy ``A[:, mask]`` with sorted indices)."""
    mask = np.asarray(mask, dtype=bool)
    newid = np.cumsum(mask) - 1
    idx = np.asarray(indices, dtype=np.int64)
    keep = mask[idx]
    ip = np.asarray(indptr, dtype=np.int64)
    rows = np.repeat(np.arange(len(ip) - 1), np.diff(ip))
    cnt = np.bincount(rows[keep], minlength=len(ip) - 1)
    new_indptr = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int64)
    return new_indptr, newid[idx[keep]].astype(np.int64), data[keep].copy()


# --------------------------------------------------------------------------------------
# randomized_svd + svd_flip (single-svdlib randomized; call sites
# pca/sparse/mod.rs:170-180,203 and pca/sparse_masked/mod.rs:341-351,364)
# --------------------------------------------------------------------------------------
def svd_flip_v(u, vt):
    """``svd_flip(Some(u), Some(vt), u_based_decision=false)``: for each component i make the
    entry of vt[i, :] with the largest magnitude positive; apply the sign to u[:, i] too."""
    j = np.argmax(np.abs(vt), axis=1)
    signs = np.sign(vt[np.arange(vt.shape[0]), j])
    signs[signs == 0] = 1
    vt = vt * signs[:, None]
    if u is not None:
        u = u * signs[None, :]
    return u, vt


def _orth(Y, normalizer):
    if normalizer == "qr":
        q, _ = np.linalg.qr(Y)
        return q
    if normalizer == "lu":
        import scipy.linalg as sla
        p, l, _ = sla.lu(Y)
        return p @ l
    if normalizer == "none":
        return Y
    raise ValueError(normalizer)


def randomized_svd(A, n_components, n_oversamples, n_power_iterations, omega,
                   mean_center=True, normalizer="qr", dtype=None):
    """Restatement of ``single_svdlib::randomized::randomized_svd`` as called at
    pca/sparse/mod.rs:170-180, taking the Gaussian test matrix
    ``omega`` (ncols x (rank+n_oversamples)) explicitly.

    A_c := A - 1 mu^T is never formed: A_c X = A X - 1 (mu^T X), A_c^T Y = A^T Y - mu (1^T Y).
    Returns (u, s, vt) truncated to rank = min(n_components, min(A.shape)), *before* svd_flip.
    """
    A = sp.csr_matrix(A)
    dtype = np.dtype(dtype or A.dtype)
    A = A.astype(dtype)
    n, m = A.shape
    rank = min(n_components, min(n, m))
    l = rank + n_oversamples
    omega = np.asarray(omega, dtype=dtype)
    assert omega.shape == (m, l), (omega.shape, (m, l))
    At = A.T.tocsr()
    mu = (np.asarray(A.sum(axis=0)).ravel().astype(np.float64) / n).astype(dtype) if mean_center \
        else np.zeros(m, dtype=dtype)

    def Ac(X):
        Y = A @ X
        if mean_center:
            Y = Y - (mu @ X)[None, :]
        return Y

    def AcT(Y):
        Z = At @ Y
        if mean_center:
            Z = Z - mu[:, None] * Y.sum(axis=0)[None, :]
        return Z

    Y = Ac(omega)
    for _ in range(n_power_iterations):
        Y = _orth(Y, normalizer)
        Z = AcT(Y)
        Z = _orth(Z, normalizer)
        Y = Ac(Z)
    Q, _ = np.linalg.qr(Y)
    B = AcT(Q).T
    Ub, s, Vt = np.linalg.svd(B, full_matrices=False)
    U = Q @ Ub
    return U[:, :rank], s[:rank], Vt[:rank, :]


def column_stats(A):
    """(sum, sumsq) over columns in f64 — what fit() derives mean_ and total_var from
    (pca/sparse/mod.rs:106-131)."""
    A = sp.csr_matrix(A)
    s = np.asarray(A.sum(axis=0)).ravel().astype(np.float64)
    A2 = A.copy()
    A2.data = A2.data.astype(np.float64) ** 2
    ss = np.asarray(A2.sum(axis=0)).ravel()
    return s, ss


class PCAResult:
    def __init__(self, components, singular_values, explained_variance, mean, total_var, u):
        self.components = components              # d x n_eff   (components_)
        self.singular_values = singular_values
        self.explained_variance = explained_variance
        self.mean = mean                          # full ncols length
        self.total_var = total_var
        self.u = u                                # n x d (discarded by the reference)


def sparse_pca_fit(A, n_components, omega=None, center=True, method="random",
                   n_oversamples=10, n_power_iterations=7, normalizer="qr", mask=None,
                   dtype=None):
    """``SparsePCA::fit`` (pca/sparse/mod.rs:102-242) and ``MaskedSparsePCA::fit``
    (pca/sparse_masked/mod.rs:255-419; pass ``mask``).

    mean_ = sum_col / n over ALL columns (:106-114 / :275-292); total_var = sum over (kept)
    columns of (sumsq - mean*sum)/(n-1) (:119-131 / :294-311); SVD on the (masked) operator;
    svd_flip with u_based_decision=false (:203 / :364); components_ = vt (:208 / :368);
    explained_variance_[i] = s[i]^2/(n-1) (:210-216 / :379-382).
    Lanczos (`method="lanczos"`) runs on the UNCENTRED operator (:134-144 / :316-331)."""
    A = sp.csr_matrix(A)
    dtype = np.dtype(dtype or A.dtype)
    n, m_full = A.shape
    s_, ss_ = column_stats(A)
    mean = (s_ / n) if center else np.zeros(m_full)
    kept = np.arange(m_full) if mask is None else np.flatnonzero(np.asarray(mask, dtype=bool))
    total_var = float(np.sum((ss_[kept] - (s_[kept] / n) * s_[kept]) / (n - 1))) if center else None
    Aop = A if mask is None else A[:, kept]
    if method == "random":
        u, s, vt = randomized_svd(Aop, n_components, n_oversamples, n_power_iterations, omega,
                                  mean_center=center, normalizer=normalizer, dtype=dtype)
    elif method == "lanczos":
        u, s, vt = truncated_svd_truth(Aop.astype(np.float64), n_components)
        u, s, vt = u.astype(dtype), s.astype(dtype), vt.astype(dtype)
    else:
        raise ValueError(method)
    u, vt = svd_flip_v(u, vt)
    ev = (s.astype(np.float64) ** 2) / (n - 1)
    if not center:
        total_var = float(ev.sum())             # pca/sparse/mod.rs:218-223
    return PCAResult(vt, s, ev.astype(dtype), mean.astype(dtype), total_var, u)


# transform modes ------------------------------------------------------------------------
EXACT = 0             # intended projection (X - 1 mu^T) V^T on the (kept) columns
REFERENCE_COMPAT = 1  # what the reference's loops actually compute


def transform(A, components, mean, center=True, mask=None, mode=EXACT):
    """``transform`` (pca/sparse/mod.rs:255-285, pca/sparse_masked/mod.rs:438-546).

    EXACT: scores = (X - 1 mu^T) V^T restricted to the kept columns.
    REFERENCE_COMPAT:
      * masked (sparse_masked/mod.rs:488-529): only STORED kept entries contribute,
        (x - mu_c) * V[k, midx(c)]; implicit zeros contribute nothing.
      * unmasked (sparse/mod.rs:268-282): the inner loop walks ``x.col_indices()`` of the whole
        matrix, so column c is visited cnt_c = nnz(column c) times and
        score[r,k] = sum_c cnt_c * (X[r,c] - mu_c) * V[k,c]."""
    A = sp.csr_matrix(A).astype(np.float64)
    V = np.asarray(components, dtype=np.float64)          # d x n_eff
    mean = np.asarray(mean, dtype=np.float64)
    kept = np.arange(A.shape[1]) if mask is None else np.flatnonzero(np.asarray(mask, dtype=bool))
    Ak = A[:, kept]
    mu = mean[kept] if center else np.zeros(len(kept))
    if mode == EXACT:
        return Ak @ V.T - (mu @ V.T)[None, :]
    if mask is not None:
        P = Ak.copy()
        P.data = np.ones_like(P.data)
        return Ak @ V.T - P @ (mu[:, None] * V.T)
    cnt = np.diff(sp.csc_matrix(A).indptr).astype(np.float64)
    Vw = V * cnt[None, :]
    return Ak @ Vw.T - (mu @ Vw.T)[None, :]


def explained_variance_ratio(ev):
    """pca/sparse/mod.rs:312-322 — normalised by the sum over the COMPUTED components."""
    ev = np.asarray(ev)
    return ev / ev.sum()


def cumulative_explained_variance_ratio(ev):
    """pca/sparse/mod.rs:333-343."""
    return np.cumsum(explained_variance_ratio(ev))


def feature_importances(components):
    """pca/sparse/mod.rs:295-302 — squared loadings."""
    return np.asarray(components) ** 2


# --------------------------------------------------------------------------------------
# truth / metrics
# --------------------------------------------------------------------------------------
def truncated_svd_truth(A, k, center_mean=None):
    """Converged top-k singular triplets — the answer ``svd_las2`` (kappa = 1e-5,
    pca/sparse/mod.rs:136-143) converges to on the UNCENTRED operator.  Dense LAPACK when
    small, ARPACK otherwise."""
    A = sp.csr_matrix(A).astype(np.float64)
    n, m = A.shape
    if n * m <= 40_000_000:
        D = A.toarray()
        if center_mean is not None:
            D = D - center_mean[None, :]
        u, s, vt = np.linalg.svd(D, full_matrices=False)
        return u[:, :k], s[:k], vt[:k]
    from scipy.sparse.linalg import svds
    u, s, vt = svds(A, k=k, tol=1e-12, random_state=0)
    o = np.argsort(-s)
    return u[:, o], s[o], vt[o]


def largest_principal_angle(V1, V2):
    """Largest principal angle (rad) between the row spaces of V1, V2 (d x n each)."""
    q1, _ = np.linalg.qr(np.asarray(V1, dtype=np.float64).T)
    q2, _ = np.linalg.qr(np.asarray(V2, dtype=np.float64).T)
    # sin-based formula is accurate for small angles
    r = q2 - q1 @ (q1.T @ q2)
    sv = np.linalg.svd(r, compute_uv=False)
    return float(np.arcsin(min(1.0, sv.max())))


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300)))
