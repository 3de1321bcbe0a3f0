#!/bin/bash
SALG_SPMM_IMPL=tm DBGS=1,31,9,17 SALG_LIB_PATH=scratch/libsalg_dbg.so timeout 300 python tools/scripts_tm_dbg.py 2>&1 | grep -E "^\[tm|^==|rror" | awk '/^==/ {print; n=0; next} {n++; if (n>4) print}'
