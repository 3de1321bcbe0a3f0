"""Probe of the fused row pass (sum_row -> normalize ROW -> log1p) on a config-5 shard: SALG_PRE_CTAS=n python tools/scripts_pre_probe.py"""
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import single_algebra_b200 as s
ctx = s.default_context()
spec = s.synth.make_spec(4_000_000, 33_000, density=0.07, seed=42)
res = []
for rep in range(3):
    d = s.synth_device(spec, 0, 500_000, dtype=np.float32, ctx=ctx)
    ctx.prof_reset(); ctx.prof_enable(True)
    d.preprocess(1e4)
    ctx.prof_enable(False)
    pr = ctx.prof()
    byt = 2 * d.nnz * 4 + (500_000 + 1) * 8
    res.append(pr["elementwise"][0])
    chk = float(np.sum(d.download_values()[:1000], dtype=np.float64))
    d.free()
print(os.environ.get("SALG_PRE_CTAS", "default"), "row pass ms", [round(x, 3) for x in res], "frac", round(byt / min(res) / 1e6 / 6451.8, 3), "chk", chk, flush=True)
