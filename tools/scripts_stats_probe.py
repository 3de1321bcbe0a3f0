import os, sys, numpy as np
sys.path.insert(0, os.getcwd())   # run from the repo root
import single_algebra_b200 as s
ctx = s.default_context()
spec = s.synth.make_spec(1_000_000, 30_000, density=0.07, seed=42)
d = s.synth_device(spec, dtype=np.float32, ctx=ctx)
mask = s.synth.make_mask(30_000, 2_000, seed=7)
om = s.synth.make_omega(2000, 60, seed=42, dtype=np.float32)
pca = s.MaskedSparsePCABuilder().n_components(50).mask(mask.tolist()).svd_method(s.SVDMethod.Random(10, 1, s.PowerIterationNormalizer.QR)).build()
for i in range(2): pca._fit(d, om, keep_scores=False, fetch=False)
ctx.prof_reset(); ctx.prof_enable(True)
for i in range(3): pca._fit(d, om, keep_scores=False, fetch=False)
ctx.sync(); ctx.prof_enable(False)
pr = ctx.prof()
print(os.environ.get("SALG_LIB_PATH", "default"), "stats ms", round(pr["stats"][0] / 3, 3), flush=True)
