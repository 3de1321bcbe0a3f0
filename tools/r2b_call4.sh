#!/bin/bash
SALG_SPMM_IMPL=tm DBGS=0,2,4,6,8,16,30 timeout 600 python tools/scripts_tm_time.py 2>&1 | grep -E "^dbg|rror" | tail -12
