#!/bin/bash
# Gram pass with 2 load stages + 3 operand buffers (experimental build) against the default 3 + 2
O=gpurun_out/r2c10
mkdir -p $O
for L in "" scratch/libsalg_gp23.so; do
if [ -n "$L" ]; then export SALG_LIB_PATH=$L; else unset SALG_LIB_PATH; fi
echo "=== lib ${L:-default}"
python - <<'PY'
import sys
sys.path.insert(0, ".")
import single_algebra_b200 as salg
ctx = salg.default_context()
for rows in (125000, 1000000):
    print("gram pass rows", rows, "ms", round(salg.op_tall_gram(ctx=ctx, device_rows=rows, k=60, iters=50)[-1], 4), flush=True)
PY
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg3 ms', round(d['ms_per_step'],3), 'gram', round(d['kernel_classes']['gram']['ms_total']/d['steps'],3))"
done
export SALG_LIB_PATH=scratch/libsalg_gp23.so
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -2
