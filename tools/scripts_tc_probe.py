import os, sys, numpy as np
sys.path.insert(0, os.getcwd())   # run from the repo root
import single_algebra_b200 as s
ctx = s.default_context()
which = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
if which == "cfg3":
    spec = s.synth.make_spec(1_000_000, 30_000, density=0.07, seed=42)
    d = s.synth_device(spec, dtype=np.float32, ctx=ctx)
    op = d.select_columns(s.synth.make_mask(30_000, 2_000, seed=7)); d.free()
else:
    spec = s.synth.make_spec(100_000, 20_000, density=0.07, seed=42)
    op = s.synth_device(spec, dtype=np.float32, ctx=ctx)
byt = op.nnz * 8 + (op.nrows + 1) * 8 + (op.ncols + op.nrows) * 60 * 4
for tr in (False, True):
    ms = s.op_spmm_bench(op, transposed=tr, k=60, iters=iters)
    print(f"{which} dbg={os.environ.get('SALG_TC_DBG','0')} transposed={tr}: {ms:.3f} ms  {byt/ms/1e6:.0f} GB/s  frac {byt/ms/1e6/6451.8:.3f}", flush=True)
