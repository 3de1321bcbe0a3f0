"""Timing decomposition of the tcgen05 product kernels (cfg3 compacted operator, one B200):
SALG_LIB_PATH=<variant .so> python tools/scripts_tc_diag.py
For each SALG_TC_DBG bit set (results are WRONG with a bit set; only the time is meaningful):
 1 no clear of the operand buffer, 2 no MMA issue, 4 no panel-slice loads, 8 no scatter stores, 16 no proxy fence."""
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
tag = os.environ.get("SALG_LIB_PATH", "default") + (" rot" if os.environ.get("SALG_TC_ROT") else "") + (" order" if os.environ.get("SALG_TC_ORDER") else "")
for dbg in [0, 1, 2, 4, 8, 16, 1 | 8, 1 | 8 | 16, 2 | 4, 1 | 2 | 8, 1 | 2 | 4 | 8 | 16]:
    os.environ["SALG_TC_DBG"] = str(dbg)
    res = []
    for tr in (False, True):
        ms = s.op_spmm_bench(op, transposed=tr, k=60, iters=8)
        res.append(f"{'AtY' if tr else 'AX'} {ms:.3f} ms ({byt/ms/1e6/6451.8:.3f})")
    print(tag, f"dbg={dbg:2d}", *res, flush=True)
