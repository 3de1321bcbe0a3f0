#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -15 > gpurun_out/t_all.log
cat gpurun_out/t_all.log | tail -5
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_cfg3.log 2> gpurun_out/bench_cfg3.err; echo "bench exit $?"
tail -c 3000 gpurun_out/bench_cfg3.log
