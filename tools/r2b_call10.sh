#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_tm.py -m gpu -q -x --timeout 300 2>&1 | tail -4
SALG_SPMM_IMPL=tm DBGS=1 SALG_LIB_PATH=scratch/libsalg_dbg.so timeout 300 python tools/scripts_tm_dbg.py 2>&1 | grep -E "^\[tm|^==|rror" | tail -4
SALG_SPMM_IMPL=tm DBGS=0,2,8,30 timeout 600 python tools/scripts_tm_time.py 2>&1 | grep -E "^dbg|rror" | tail -12
