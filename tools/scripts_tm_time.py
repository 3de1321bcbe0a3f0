"""Stand-alone timing of the two products at the config-3 operator (1M x 2000 kept genes): python tools/scripts_tm_time.py"""
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import single_algebra_b200._native as N
if os.environ.get("SALG_LIB_PATH"):
    N.LIB_PATH = os.path.abspath(os.environ["SALG_LIB_PATH"])
import single_algebra_b200 as s
ctx = s.default_context()
spec = s.synth.make_spec(1_000_000, 30_000, density=0.07, seed=42)
d = s.synth_device(spec, dtype=np.float32, ctx=ctx)
op = d.select_columns(s.synth.make_mask(30_000, 2_000, seed=7)); d.free()
byt = op.nnz * 8 + (op.nrows + 1) * 8 + (op.ncols + op.nrows) * 60 * 4
for dbg in os.environ.get("DBGS", "0").split(","):
    os.environ["SALG_TM_DBG"] = dbg
    res = []
    for tr in (False, True):
        ms = s.op_spmm_bench(op, transposed=tr, k=60, iters=10)
        res.append(f"{'AtY' if tr else 'AX'} {ms:.3f} ms ({byt/ms/1e6/6451.8:.3f})")
    print("dbg", dbg, *res, flush=True)
