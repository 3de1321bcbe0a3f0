import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import single_algebra_b200 as s
ctx = s.default_context()
Y = np.random.default_rng(0).standard_normal((4000, 60)).astype(np.float32)
for _ in range(3):
    q, r = s.op_cholqr2(Y, ctx)
print("ok", np.abs(q.T @ q - np.eye(60)).max())
