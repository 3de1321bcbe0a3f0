#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/b15.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches15.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu15.log 2>&1; echo "ncu exit $?"
