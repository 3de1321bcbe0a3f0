#!/bin/bash
for d in 8 15; do
SALG_GP_DBG=$d python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import single_algebra_b200 as s
g, cs, ms = s.op_tall_gram(None, device_rows=1_000_000, k=60, iters=5)
print("dbg", os.environ.get("SALG_GP_DBG"), f"{ms:.4f} ms", flush=True)
PY
done
