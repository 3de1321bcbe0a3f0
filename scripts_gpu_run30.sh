#!/bin/bash
mkdir -p gpurun_out
SALG_TC_ORDER=1 timeout 900 ncu --set full --clock-control none -k regex:"tc_ax_kernel|tc_aty_kernel" -s 5 -c 2 -o gpurun_out/prof_tc4 -f python scripts_tc_probe.py cfg3 1 > gpurun_out/ncu_tc4.log 2>&1; echo "ncu exit $?"
