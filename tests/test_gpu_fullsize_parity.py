"""Full-size ORACLE parity (not just properties) for BASELINE configs 2 and 3: the CUDA fit against the multi-threaded
C++ restatement of the reference path (oracle/cpu_ref.cpp, f64 arithmetic, same host-generated Omega) on the FULL operator.
Tolerances (north-star): singular values 1e-4 relative for an f32 fit against the f64 oracle, largest principal angle
< 1e-3 rad, column statistics 1e-4 (f32).  The oracle takes ~1 minute per config on the GPU box's 16 host cores."""
import numpy as np
import pytest

from oracle import cpu_ref as R
from oracle import oracle as O

pytestmark = pytest.mark.gpu

SIGMA_TOL_F32 = 1e-4
ANGLE_TOL = 1e-3


def test_config2_full_operator_parity(salg, ctx):
    """SparsePCA f32, 100k x 20k @7 %, Random{10, 7, QR}, k = 50 — every stored entry of the operator takes part."""
    spec = salg.synth.make_spec(100_000, 20_000, density=0.07, seed=42)
    d = salg.synth_device(spec, dtype=np.float32, ctx=ctx)
    off = np.empty(d.nrows + 1, np.int64); idx = np.empty(d.nnz, np.uint32); val = np.empty(d.nnz, np.float32)
    d.download_raw(off, idx, val)
    om = salg.synth.make_omega(20_000, 60, seed=42, dtype=np.float32)
    pca = salg.SparsePCABuilder().n_components(50).svd_method(
        salg.SVDMethod.Random(10, 7, salg.PowerIterationNormalizer.QR)).build()
    scores = pca.fit_transform(d, omega=om)
    R.set_threads()
    ref = R.pca_fit(off, idx, val.astype(np.float64), 100_000, 20_000, 50, om.astype(np.float64))
    assert O.rel_err(pca.singular_values_, ref.singular_values) < SIGMA_TOL_F32
    assert O.largest_principal_angle(pca.components_, ref.components) < ANGLE_TOL
    assert np.allclose(pca.mean_, ref.mean, rtol=1e-4, atol=1e-7)
    assert abs(pca.total_var_ - ref.total_var) < 1e-4 * ref.total_var
    assert np.allclose(pca.explained_variance_, ref.explained_variance, rtol=2e-4)
    # the projection of the fitted rows, entry by entry (same signs: svd_flip on both sides)
    assert np.abs(scores - ref.scores).max() < 2e-3 * np.abs(ref.scores).max()
    d.free()


def test_config3_full_operator_parity(salg, ctx):
    """MaskedSparsePCA f32, 1M x 30k @7 %, 2000-gene mask, Random{10, 7, QR}, k = 50: the fused statistics + compaction +
    tile path on the full 2.1e9-entry matrix against the oracle on the compacted 1M x 2000 operator (what
    MaskedCSRMatrix presents to the SVD engine, bit-exact per test_gpu_csr)."""
    spec = salg.synth.make_spec(1_000_000, 30_000, density=0.07, seed=42)
    d = salg.synth_device(spec, dtype=np.float32, ctx=ctx)
    mask = salg.synth.make_mask(30_000, 2_000, seed=7)
    om = salg.synth.make_omega(2_000, 60, seed=42, dtype=np.float32)
    pca = salg.MaskedSparsePCABuilder().n_components(50).mask(mask.tolist()).svd_method(
        salg.SVDMethod.Random(10, 7, salg.PowerIterationNormalizer.QR)).build()
    scores = pca.fit_transform(d, omega=om)
    s_all, q_all, _, _ = d.col_stats()                       # f64 column statistics of the full matrix
    op = d.select_columns(mask)
    d.free()
    off = np.empty(op.nrows + 1, np.int64); idx = np.empty(op.nnz, np.uint32); val = np.empty(op.nnz, np.float32)
    op.download_raw(off, idx, val)
    op.free()
    R.set_threads()
    ref = R.pca_fit(off, idx, val.astype(np.float64), 1_000_000, 2_000, 50, om.astype(np.float64))
    assert O.rel_err(pca.singular_values_, ref.singular_values) < SIGMA_TOL_F32
    assert O.largest_principal_angle(pca.components_, ref.components) < ANGLE_TOL
    # mean_ has the FULL column count (pca/sparse_masked/mod.rs:280-291); its kept entries are the operator's means
    assert pca.mean_.shape == (30_000,)
    assert np.allclose(pca.mean_[mask], ref.mean, rtol=1e-4, atol=1e-7)
    assert np.allclose(pca.mean_, s_all / 1e6, rtol=1e-4, atol=1e-7)
    assert abs(pca.total_var_ - ref.total_var) < 1e-4 * ref.total_var
    assert np.abs(scores - ref.scores).max() < 2e-3 * np.abs(ref.scores).max()
