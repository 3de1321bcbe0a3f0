#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_tm.py -m gpu -q -x --timeout 300 2>&1 | tail -4
SALG_SPMM_IMPL=tm SALG_TM_DBG=1 SALG_LIB_PATH=scratch/libsalg_dbg.so timeout 300 python tools/scripts_tc_probe2.py 2>&1 | grep -E "^\[tm|adjoint" | tail -28 | awk "NR<=4 || NR>24"
SALG_SPMM_IMPL=tm DBGS=0,32,8,30 timeout 600 python tools/scripts_tm_time.py 2>&1 | grep -E "^dbg|rror" | tail -12
