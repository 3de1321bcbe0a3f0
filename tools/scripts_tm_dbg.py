"""Cycle counters of the TMEM-operand products (experiment build, -DTM_DBG_=1): DBGS=1,31 SALG_LIB_PATH=... python tools/scripts_tm_dbg.py"""
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
os.environ["SALG_TM_DBG"] = "0"
for tr in (False, True):
    s.op_spmm_bench(op, transposed=tr, k=60, iters=2)
for dbg in os.environ.get("DBGS", "1").split(","):
    os.environ["SALG_TM_DBG"] = dbg
    print("== dbg", dbg, flush=True)
    for tr in (False, True):
        s.op_spmm_bench(op, transposed=tr, k=60, iters=1)
