#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 200 -x -k "tc_products or spmm" 2>&1 | tail -15
SALG_TC_DBG=32 timeout 300 python scripts_tc_probe.py cfg3 10 2>&1 | grep -v Warn | grep "mma total\|cfg3" | tail -3 | tee gpurun_out/tc_probe.log
timeout 300 python scripts_tc_probe.py cfg2 10 2>&1 | grep -v Warn | tee -a gpurun_out/tc_probe.log
