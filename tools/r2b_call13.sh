#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_tm.py -m gpu -q -x --timeout 300 2>&1 | tail -3
SALG_SPMM_IMPL=tm DBGS=0,64,0 timeout 600 python tools/scripts_tm_time.py 2>&1 | grep -E "^dbg|rror" | tail -12
