"""Probe (GPU): average time of the fused Gram + pre-split pass (tc_gram_prep_kernel) for a 125k-row (8-GPU shard) and a 1M-row panel
under its timing switches SALG_GP_DBG (1 no store, 2 no MMA, 4 no load, 32 no Gram; results are wrong with a bit set).
profiles/r02c_summary.md section 3."""
import os, sys, subprocess
CODE = r'''
import sys, numpy as np
sys.path.insert(0, ".")
import single_algebra_b200 as salg
ctx = salg.default_context()
for rows in (125000, 1000000):
    r = salg.op_tall_gram(ctx=ctx, device_rows=rows, k=60, iters=50)
    print("rows", rows, "ms", r if not isinstance(r, tuple) else r[-1], flush=True)
'''
for dbg in ("0", "7", "39"):
    e = dict(os.environ); e["SALG_GP_DBG"] = dbg
    print("=== SALG_GP_DBG", dbg, flush=True)
    subprocess.run([sys.executable, "-c", CODE], env=e)
