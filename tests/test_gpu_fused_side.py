"""Round-2 fusions of the replicated small side (csrc/dense.cu: zside_solve_kernel, csrc/tm.cu: tm_zside_apply_kernel) and of
the tail of fit_transform (scores = (A_c Q_B) U_R, the big product beside the Jacobi SVD): the fused schedule must reproduce the
explicit chain — chol_inv, panel_mul, panel_gram, chol_inv, panel_mul, column sums, |max|, pre-split, then SVD, then A_c V —
which stays selectable (SALG_NO_ZSIDE / SALG_NO_TAIL_OVERLAP are read per fit) and is the schedule the round-1 parity tests pinned.
Reference: single-svdlib randomized_svd as called at /root/reference/src/dimred/pca/sparse/mod.rs:170-180 and
src/dimred/pca/sparse_masked/mod.rs:341-351 (QR normaliser after every product)."""
import os

import numpy as np
import pytest

from conftest import planted_counts
from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _fit(salg, ctx, A, mask, n_eff, env):
    old = {k: os.environ.get(k) for k in ("SALG_NO_ZSIDE", "SALG_NO_TAIL_OVERLAP")}
    for k in old:
        os.environ.pop(k, None)
    os.environ.update(env)
    try:
        om = salg.synth.make_omega(n_eff, 40, seed=42, dtype=np.float32)
        rnd = salg.SVDMethod.Random(10, 7, salg.PowerIterationNormalizer.QR)
        if mask is not None:
            pca = salg.MaskedSparsePCABuilder().n_components(30).mask(mask.tolist()).svd_method(rnd).build()
        else:
            pca = salg.SparsePCABuilder().n_components(30).svd_method(rnd).build()
        scores = pca.fit_transform(salg.CsrMatrix.from_scipy(A, ctx), omega=om)
        return pca, scores
    finally:
        for k, v in old.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v


@pytest.mark.parametrize("masked", [False, True])
def test_fused_small_side_matches_explicit_chain(salg, ctx, masked):
    ctx.set_spmm_impl("tm")
    A = planted_counts(20_000, 3000, seed=11, dtype=np.float32)
    mask = salg.synth.make_mask(3000, 700, seed=7) if masked else None
    n_eff = 700 if masked else 3000
    fused, s_f = _fit(salg, ctx, A, mask, n_eff, {})
    plain, s_p = _fit(salg, ctx, A, mask, n_eff, {"SALG_NO_ZSIDE": "1", "SALG_NO_TAIL_OVERLAP": "1"})
    assert fused.numeric_flags() == 0 and plain.numeric_flags() == 0
    assert O.rel_err(fused.singular_values_, plain.singular_values_) < 5e-5
    assert O.largest_principal_angle(fused.components_, plain.components_) < 5e-4
    # scores: (A_c Q_B) U_R diag(sign) against A_c V of the explicit tail, and against the exact projection
    assert np.abs(s_f - s_p).max() < 2e-4 * np.abs(s_p).max()
    ex = O.transform(A, fused.components_, fused.mean_, center=True, mask=mask, mode=O.EXACT)
    assert np.abs(s_f - ex).max() < 2e-4 * np.abs(ex).max()
    ref = O.sparse_pca_fit(A.astype(np.float64), 30, omega=salg.synth.make_omega(n_eff, 40, seed=42, dtype=np.float32).astype(np.float64),
                           mask=mask, n_oversamples=10, n_power_iterations=7)
    assert O.rel_err(fused.singular_values_, ref.singular_values) < 1e-4
    assert O.largest_principal_angle(fused.components_, ref.components) < 1e-3


def test_rank_deficient_sketch_drops_dependent_columns(salg, ctx):
    """Fewer independent directions than sketch columns (3000 cells that are copies of 12 profiles: centred rank 11, sketch
    l = 30).  CholeskyQR marks the dependent columns (numeric flag bit 1): a first pass floors their pivots without letting
    them update the trailing matrix, the last pass drops what is still dependent (zero columns of Q) — in the fused kernels
    exactly as in the explicit chain.  The 11 real triplets agree with the dense SVD and the rest are (near) zero.
    Regression: rounds 1-2 floored every pivot and kept the sub-diagonal columns — this input returned non-finite singular
    values (f32) or a spurious one (f64); the Jacobi SVD produced NaN from columns 1e-17 apart in norm.
    Without power iterations the amplified noise columns of the FIRST tall panel carry the all-ones direction, whose implicit
    centring term cancels catastrophically: that case must fail loudly (Frobenius-norm check), never return a wrong value."""
    rng = np.random.default_rng(3)
    import scipy.sparse as sp
    base = (rng.random((12, 400)) < 0.2) * rng.integers(1, 6, size=(12, 400))
    D = base[rng.integers(0, 12, size=3000)]
    s_true = np.linalg.svd(D.astype(np.float64) - D.astype(np.float64).mean(axis=0), compute_uv=False)
    old = os.environ.pop("SALG_NO_ZSIDE", None)
    try:
        for dtype, tol in ((np.float32, 1e-4), (np.float64, 1e-9)):
            A = sp.csr_matrix(D.astype(dtype))
            for env in ({}, {"SALG_NO_ZSIDE": "1"}):
                for q in (0, 4):
                    os.environ.pop("SALG_NO_ZSIDE", None)
                    os.environ.update(env)
                    om = salg.synth.make_omega(400, 30, seed=42, dtype=dtype)
                    pca = salg.SparsePCABuilder().n_components(20).svd_method(
                        salg.SVDMethod.Random(10, q, salg.PowerIterationNormalizer.QR)).build()
                    try:
                        pca.fit(salg.CsrMatrix.from_scipy(A, ctx), omega=om)
                    except salg.SalgError as e:
                        assert q == 0 and "Frobenius" in str(e), (env, q, dtype, str(e))
                        continue
                    s = pca.singular_values_
                    V = pca.components_.astype(np.float64)
                    assert np.all(np.isfinite(s)) and np.all(np.isfinite(V)), (env, q, dtype)
                    assert pca.numeric_flags() & 1
                    assert O.rel_err(s[:11], s_true[:11]) < tol, (env, q, dtype, s[:12], s_true[:12])
                    assert np.all(s[11:] < 1e-3 * s[0]), (env, q, dtype, s[11:])
                    assert np.abs(V[:11] @ V[:11].T - np.eye(11)).max() < 1e-4
    finally:
        os.environ.pop("SALG_NO_ZSIDE", None)
        if old is not None:
            os.environ["SALG_NO_ZSIDE"] = old
