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
os.environ["SALG_TC_PFD"] = "0"
tag = os.environ.get("SALG_LIB_PATH", "default")
for dbg in (0, 2, 127, 127 | 128, 2 | 128, 32 | 127 | 128, 32 | 2 | 128):
    os.environ["SALG_TC_DBG"] = str(dbg)
    ms = s.op_spmm_bench(op, transposed=False, k=60, iters=(1 if dbg & 32 else 8))
    print(f"{tag} dbg={dbg} AX {ms:.3f} ms", file=sys.stderr, flush=True)
