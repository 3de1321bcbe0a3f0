"""A/B probe of the statistics kernels in experiment builds of the library:
SALG_LIB_PATH=scratch/libsalg_x.so python tools/scripts_stats_probe.py
Prints the PROF_STATS class time (CUDA events on the library's stream) of: the masked statistics + fused compaction pass
of config 3, the unmasked column statistics of a config-5 shard (33k columns: column-tiled kernel) on raw counts, and of
config 2 (20k columns: flat kernel), with a checksum of each result so a broken variant is not mistaken for a fast one."""
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())   # run from the repo root
import single_algebra_b200._native as N
if os.environ.get("SALG_LIB_PATH"):
    N.LIB_PATH = os.path.abspath(os.environ["SALG_LIB_PATH"])
import single_algebra_b200 as s

ctx = s.default_context()


def stats_ms(fn, reps=3):
    fn()
    ctx.prof_reset(); ctx.prof_enable(True)
    for _ in range(reps):
        r = fn()
    ctx.prof_enable(False)
    return ctx.prof()["stats"][0] / reps, r


out = [os.environ.get("SALG_LIB_PATH", "default")]
# config 3: masked fit (statistics class = masked statistics + fused compaction)
spec = s.synth.make_spec(1_000_000, 30_000, density=0.07, seed=42)
d = s.synth_device(spec, dtype=np.float32, ctx=ctx)
mask = s.synth.make_mask(30_000, 2_000, seed=7)
om = s.synth.make_omega(2_000, 60, seed=42, dtype=np.float32)


def fit():
    pca = s.MaskedSparsePCABuilder().n_components(50).mask(mask.tolist()).svd_method(
        s.SVDMethod.Random(10, 7, s.PowerIterationNormalizer.QR)).build()
    pca.fit(d, omega=om)
    return pca


ms, pca = stats_ms(fit)
out.append(f"cfg3 masked stats {ms:.3f} ms (s1 {pca.singular_values_[0]:.6e}, mean sum {float(np.sum(pca.mean_, dtype=np.float64)):.9e})")
d.free()
# config 5 shard: unmasked, 33k columns -> column-tiled kernel
spec = s.synth.make_spec(4_000_000, 33_000, density=0.07, seed=42)
d = s.synth_device(spec, 0, 500_000, dtype=np.float32, ctx=ctx)
ms, r = stats_ms(lambda: d.sum_col_and_squared())
out.append(f"cfg5 tiled stats {ms:.3f} ms (sum {float(np.sum(r[0], dtype=np.float64)):.9e}, sumsq {float(np.sum(r[1], dtype=np.float64)):.9e})")
d.free()
# config 2: unmasked, 20k columns -> flat kernel
spec = s.synth.make_spec(100_000, 20_000, density=0.07, seed=42)
d = s.synth_device(spec, dtype=np.float32, ctx=ctx)
ms, r = stats_ms(lambda: d.sum_col_and_squared())
out.append(f"cfg2 flat stats {ms:.3f} ms (sum {float(np.sum(r[0], dtype=np.float64)):.9e}, sumsq {float(np.sum(r[1], dtype=np.float64)):.9e})")
d.free()
print(" | ".join(out), flush=True)
