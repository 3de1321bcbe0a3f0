"""A/B probe of experiment builds of the library: SALG_LIB_PATH=scratch/libsalg_x.so [SALG_TC_ORDER=1] python scripts_tc_probe2.py"""
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())   # run from the repo root
import single_algebra_b200._native as N
if os.environ.get("SALG_LIB_PATH"):
    N.LIB_PATH = os.path.abspath(os.environ["SALG_LIB_PATH"])
import single_algebra_b200 as s
ctx = s.default_context()
spec = s.synth.make_spec(1_000_000, 30_000, density=0.07, seed=42)
d = s.synth_device(spec, dtype=np.float32, ctx=ctx)
op = d.select_columns(s.synth.make_mask(30_000, 2_000, seed=7)); d.free()
byt = op.nnz * 8 + (op.nrows + 1) * 8 + (op.ncols + op.nrows) * 60 * 4
# correctness of the variant on the fly: A^T (A X) against the default library is covered by the tests; here a cheap
# self-check (adjoint identity) so a broken variant is not mistaken for a fast one
rng = np.random.default_rng(0)
X = rng.standard_normal((op.ncols, 60)).astype(np.float32)
Y = rng.standard_normal((op.nrows, 60)).astype(np.float32)
AX = s.op_spmm(op, X).astype(np.float64)
AtY = s.op_spmm(op, Y, transposed=True).astype(np.float64)
lhs, rhs = np.sum(AX * Y, dtype=np.float64), np.sum(X * AtY, dtype=np.float64)
ok = abs(lhs - rhs) < 1e-5 * np.sqrt(np.sum(AX ** 2) * np.sum(Y.astype(np.float64) ** 2))
res = []
for tr in (False, True):
    ms = s.op_spmm_bench(op, transposed=tr, k=60, iters=10)
    res.append(f"{'AtY' if tr else 'AX'} {ms:.3f} ms ({byt/ms/1e6/6451.8:.3f})")
print(os.environ.get("SALG_LIB_PATH", "default"), "order" if os.environ.get("SALG_TC_ORDER") else "", "adjoint-ok" if ok else "ADJOINT-FAIL", *res, flush=True)
