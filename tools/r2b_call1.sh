#!/bin/bash
# first GPU run of the TMEM-operand products: parity tests, then the stand-alone product timings of both generations
O=gpurun_out/r2b1
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_tm.py -m gpu -q -x --timeout 300 2>&1 | tail -25
SALG_SPMM_IMPL=tc timeout 300 python tools/scripts_tc_probe2.py 2>&1 | tail -3
SALG_SPMM_IMPL=tm timeout 300 python tools/scripts_tc_probe2.py 2>&1 | tail -3
