"""tm vs tc products on device-generated operators: SHAPES=rows:cols,... python tools/scripts_tm_vs_tc.py"""
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import single_algebra_b200 as s
ctx = s.default_context()
rng = np.random.default_rng(0)
for shp in os.environ.get("SHAPES", "100000:20000").split(","):
    nr, nc = map(int, shp.split(":"))
    spec = s.synth.make_spec(nr, nc, density=0.07, seed=42)
    d = s.synth_device(spec, dtype=np.float32, ctx=ctx)
    for tr in (False, True):
        X = rng.standard_normal((nr if tr else nc, 60)).astype(np.float32)
        ctx.set_spmm_impl("tc"); a = s.op_spmm(d, X, transposed=tr)
        ctx.set_spmm_impl("tm"); b = s.op_spmm(d, X, transposed=tr)
        err = np.abs(a - b)
        bad = np.argwhere(err > 1e-4 * np.abs(a).max())
        print(shp, "AtY" if tr else "AX", "max rel err", err.max() / np.abs(a).max(), "bad entries", len(bad),
              "rows" if len(bad) else "", np.unique(bad[:, 0])[:12] if len(bad) else "", np.unique(bad[:, 0] // 128)[:12] if len(bad) else "", flush=True)
    d.free()
