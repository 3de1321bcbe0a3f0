#!/bin/bash
mkdir -p gpurun_out
for f in csr stats ops pca scale; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -m gpu -q --timeout 600 > gpurun_out/t_$f.log 2>&1
  echo "$f exit $?"; tail -4 gpurun_out/t_$f.log
done
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_cfg3.log 2>&1; echo "bench cfg3 exit $?"; tail -c 2600 gpurun_out/bench_cfg3.log
timeout 900 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_cfg2.log 2>&1; echo "bench cfg2 exit $?"; tail -c 2000 gpurun_out/bench_cfg2.log
