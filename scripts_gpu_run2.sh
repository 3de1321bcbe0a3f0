#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_scale.py -m gpu -q --timeout 900 > gpurun_out/t_scale.log 2>&1; echo "scale exit $?"; tail -15 gpurun_out/t_scale.log
timeout 1200 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_cfg3.log 2>&1; echo "bench cfg3 exit $?"; tail -c 3000 gpurun_out/bench_cfg3.log
timeout 900 python bench.py --workload cfg2 --steps 5 --warmup 3 > gpurun_out/bench_cfg2.log 2>&1; echo "bench cfg2 exit $?"; tail -c 2500 gpurun_out/bench_cfg2.log
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench ref exit $?"; tail -c 1200 gpurun_out/bench_ref.log
timeout 900 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/plain.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_cfg3.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu.log 2>&1; echo "ncu exit $?"; tail -3 gpurun_out/ncu.log
