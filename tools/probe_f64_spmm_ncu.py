"""Probe (GPU, run under ncu): the f64 sparse x panel products (spmm_chunk_kernel, CUDA cores) on the reference's own test shape scaled
to 1M rows: 1M x 2500 at 1 % density (25 entries per row), panel of 60 columns.  profiles/r02c_summary.md section 5."""
import sys
import numpy as np
sys.path.insert(0, ".")
import single_algebra_b200 as salg
ctx = salg.default_context()
spec = salg.synth.make_spec(1_000_000, 2500, density=0.01, seed=42)
d = salg.synth_device(spec, dtype=np.float64, ctx=ctx)
rng = np.random.default_rng(0)
X = rng.standard_normal((2500, 60))
Y = rng.standard_normal((1_000_000, 60))
print("nnz", d.nnz, flush=True)
for it in range(2):
    a = salg.op_spmm(d, X)
    b = salg.op_spmm(d, Y, transposed=True)
print("done", a.shape, b.shape, flush=True)
