#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_gram_prep -s 2 -c 1 -o gpurun_out/prof_gp -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_gp.log 2>&1; echo "ncu exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:col_stats_masked -s 1 -c 1 -o gpurun_out/prof_stats -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_stats.log 2>&1; echo "ncu exit $?"
