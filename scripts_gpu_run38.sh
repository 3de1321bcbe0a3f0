#!/bin/bash
echo "== s640"; SALG_LIB_PATH=scratch/libsalg_s640.so timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q --timeout 300 2>&1 | tail -4
echo "== g4"; SALG_LIB_PATH=scratch/libsalg_g4.so timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q --timeout 300 2>&1 | grep -E "^tests|passed|failed|Error" | tail -6
SALG_LIB_PATH=scratch/libsalg_g4.so python scripts_tc_probe2.py 2>&1 | tail -1
