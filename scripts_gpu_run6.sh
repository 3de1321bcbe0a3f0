#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 200 -x > gpurun_out/t_ops.log 2>&1; echo "ops exit $?"; tail -3 gpurun_out/t_ops.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_cfg3.log 2>&1; echo "bench cfg3 exit $?"; tail -c 2800 gpurun_out/bench_cfg3.log
