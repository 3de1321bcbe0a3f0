"""Cycle counters of CTA 0 (scatter group 0, MMA thread, epilogue warp 0, entry loader 0) in the product kernels."""
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import single_algebra_b200 as s
ctx = s.default_context()
spec = s.synth.make_spec(1_000_000, 30_000, density=0.07, seed=42)
d = s.synth_device(spec, dtype=np.float32, ctx=ctx)
op = d.select_columns(s.synth.make_mask(30_000, 2_000, seed=7)); d.free()
os.environ["SALG_TC_PFD"] = "0"
for dbg in (32, 32 | 2, 32 | 31, 32 | 95):
    os.environ["SALG_TC_DBG"] = str(dbg)
    print(f"==== dbg={dbg}", file=sys.stderr, flush=True)
    for tr in (False, True):
        ms = s.op_spmm_bench(op, transposed=tr, k=60, iters=1)
        print(f"dbg={dbg} {'AtY' if tr else 'AX'} {ms:.3f} ms", file=sys.stderr, flush=True)
