#!/bin/bash
for d in 0 16 7 23; do
SALG_GP_DBG=$d python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import single_algebra_b200 as s
g, cs, ms = s.op_tall_gram(None, device_rows=1_000_000, k=60, iters=10)
print("dbg", os.environ.get("SALG_GP_DBG"), f"{ms:.4f} ms", flush=True)
PY
done
timeout 600 python -m pytest tests/test_gpu_pca.py -m gpu -x -q --timeout 300 -k "power_iteration or golden" 2>&1 | tail -2
