"""Probe (GPU): randomized fits of an exactly rank-deficient matrix (3000 cells = copies of 12 profiles, sketch l = 30) under the
schedule switches of the library; prints the leading singular values against the dense SVD.  See DESIGN.md 4b."""
import os, sys, subprocess
CODE = r'''
import os, numpy as np, scipy.sparse as sp, sys
sys.path.insert(0, ".")
import single_algebra_b200 as salg
ctx = salg.default_context()
impl = os.environ.get("DBG_IMPL")
if impl: ctx.set_spmm_impl(impl)
dt = np.float64 if os.environ.get("DBG_F64") else np.float32
rng = np.random.default_rng(3)
base = (rng.random((12, 400)) < 0.2) * rng.integers(1, 6, size=(12, 400))
D = base[rng.integers(0, 12, size=3000)].astype(dt)
A = sp.csr_matrix(D)
s_true = np.linalg.svd(D.astype(np.float64) - D.astype(np.float64).mean(axis=0), compute_uv=False)
for q in (0, 2):
    om = salg.synth.make_omega(400, 30, seed=42, dtype=dt)
    pca = salg.SparsePCABuilder().n_components(20).svd_method(salg.SVDMethod.Random(10, q, salg.PowerIterationNormalizer.QR)).build()
    try:
        pca.fit(salg.CsrMatrix.from_scipy(A, ctx), omega=om)
        s = pca.singular_values_
        print("q", q, "ok flags", pca.numeric_flags(), "s[:3]", s[:3], "true", s_true[:3], "s[9:13]", s[9:13], "true", s_true[9:13], flush=True)
    except Exception as e:
        print("q", q, "FAILED", e, flush=True)
'''
for env in ({}, {"SALG_NO_ZSIDE": "1"}, {"SALG_NO_FUSED_FINAL": "1"}, {"DBG_IMPL": "chunk"}, {"DBG_F64": "1"}):
    e = dict(os.environ); e.update(env)
    print("=== env", env, flush=True)
    subprocess.run([sys.executable, "-c", CODE], env=e)
