#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts_tc_probe.py cfg3 1 > gpurun_out/probe27.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_ax_kernel|tc_aty_kernel" -s 6 -c 2 -o gpurun_out/prof_tc3 -f python scripts_tc_probe.py cfg3 1 > gpurun_out/ncu_tc3.log 2>&1; echo "ncu exit $?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/b27.log 2> gpurun_out/b27.err; echo "bench exit $?"
tail -c 600 gpurun_out/b27.log
