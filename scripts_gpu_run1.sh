#!/bin/bash
# first GPU pass: environment probe, per-file GPU tests (separate processes), smoke, tiny bench
mkdir -p gpurun_out
(free -g; nproc; nvidia-smi -L; nvidia-smi --query-gpu=memory.total,clocks.max.sm --format=csv) > gpurun_out/host.txt 2>&1
for f in csr stats ops pca; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -m gpu -q --timeout 600 > gpurun_out/t_$f.log 2>&1
  echo "exit $?" >> gpurun_out/t_$f.log
  tail -5 gpurun_out/t_$f.log
done
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --workload tiny --steps 3 --warmup 3 --cpu-sample-rows 5000 > gpurun_out/bench_tiny.log 2>&1; echo "bench tiny exit $?"; tail -c 1500 gpurun_out/bench_tiny.log
