"""Entry-stream experiments on the tcgen05 product kernels (cfg3 compacted operator, one B200):
L2 prefetch distance (SALG_TC_PFD, units) x timing switches (SALG_TC_DBG: 31 = skeleton only, 95 = skeleton without
entry loads; results are wrong with a switch set).  SALG_LIB_PATH selects a build variant (ring depth)."""
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
tag = os.environ.get("SALG_LIB_PATH", "default")
# correctness of the variant with prefetch on (adjoint identity)
os.environ["SALG_TC_DBG"] = "0"; os.environ["SALG_TC_PFD"] = "32"
rng = np.random.default_rng(0)
X = rng.standard_normal((op.ncols, 60)).astype(np.float32)
Y = rng.standard_normal((op.nrows, 60)).astype(np.float32)
AX = s.op_spmm(op, X).astype(np.float64)
AtY = s.op_spmm(op, Y, transposed=True).astype(np.float64)
lhs, rhs = np.sum(AX * Y, dtype=np.float64), np.sum(X * AtY, dtype=np.float64)
print(tag, "adjoint", abs(lhs - rhs) / np.sqrt(np.sum(AX ** 2) * np.sum(Y.astype(np.float64) ** 2)), flush=True)
for pfd in (0, 8, 16, 32, 64, 128):
    os.environ["SALG_TC_PFD"] = str(pfd)
    for dbg in (0, 2, 31, 95):
        os.environ["SALG_TC_DBG"] = str(dbg)
        res = []
        for tr in (False, True):
            ms = s.op_spmm_bench(op, transposed=tr, k=60, iters=8)
            res.append(f"{'AtY' if tr else 'AX'} {ms:.3f} ms ({byt/ms/1e6/6451.8:.3f})")
        print(tag, f"pfd={pfd:3d} dbg={dbg:2d}", *res, flush=True)
