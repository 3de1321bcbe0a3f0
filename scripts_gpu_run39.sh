#!/bin/bash
for v in a b c; do
echo "== $v"; SALG_LIB_PATH=scratch/libsalg_$v.so timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q --timeout 300 2>&1 | grep -E "passed|failed|Error" | tail -3
SALG_LIB_PATH=scratch/libsalg_$v.so python scripts_tc_probe2.py 2>&1 | tail -1
done
