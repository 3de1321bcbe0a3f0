#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q --timeout 300 -k tall_gram 2>&1 | tail -12
for d in 0 1 2 4 3 7; do
SALG_GP_DBG=$d python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import single_algebra_b200 as s
g, cs, ms = s.op_tall_gram(None, device_rows=1_000_000, k=60, iters=10)
print("dbg", os.environ.get("SALG_GP_DBG"), f"{ms:.4f} ms", flush=True)
PY
done
