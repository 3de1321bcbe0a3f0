"""A small masked randomized fit (4000 x 1500, 400 kept genes, k = 50: the same 60 x 60 small-side problem as the benchmark
configs) — ncu target for the replicated small-side kernels (jacobi_svd64_kernel, chol_inv_kernel)."""
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import single_algebra_b200 as s
ctx = s.default_context()
spec = s.synth.make_spec(4000, 1500, density=0.07, seed=42)
dev = s.synth_device(spec, dtype=np.float32, ctx=ctx)
mask = s.synth.make_mask(1500, 400, seed=7)
om = s.synth.make_omega(400, 60, seed=42, dtype=np.float32)
for _ in range(3):
    pca = s.MaskedSparsePCABuilder().n_components(50).mask(mask.tolist()).svd_method(
        s.SVDMethod.Random(10, 7, s.PowerIterationNormalizer.QR)).build()
    pca.fit(dev, omega=om)
print("sigma0", pca.singular_values_[0], "flags", pca.numeric_flags())
