#!/bin/bash
# cycle counters of the TMEM-operand products (experiment build) + ncu --set full of both kernels
O=gpurun_out/r2b2
mkdir -p $O
SALG_SPMM_IMPL=tm SALG_TM_DBG=1 SALG_LIB_PATH=scratch/libsalg_dbg.so timeout 300 python tools/scripts_tc_probe2.py 2>&1 | grep -E "^\[tm|adjoint" | tail -8
SALG_SPMM_IMPL=tm timeout 300 python tools/scripts_tc_probe2.py > $O/plain.log 2>&1 &&
SALG_SPMM_IMPL=tm timeout 900 ncu --set full --clock-control none --import-source on -k regex:tm_product -s 2 -c 2 -o $O/tm_products python tools/scripts_tc_probe2.py > $O/ncu.log 2>&1
tail -3 $O/ncu.log
