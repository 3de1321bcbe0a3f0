#!/bin/bash
O=gpurun_out/r2c11
mkdir -p $O
for cfg in "10 4e-6" "10 2e-5" "10 1e-4" "0 4e-6" "4 1e-4" "6 1e-3"; do
set -- $cfg
SALG_JACOBI_DBG=1 SALG_JACOBI_F32=$1 SALG_JACOBI_TOL32=$2 timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > $O/b_$1_$2.json 2> $O/b_$1_$2.err
echo "f32=$1 tol=$2: $(tail -1 $O/b_$1_$2.err) $(python -c "
import json; d=json.load(open('$O/b_$1_$2.json')); print(round(d['ms_per_step'],2), round(d['kernel_classes']['jacobi']['ms_total']/d['steps'],3))")"
done
timeout 600 python -m pytest tests/test_gpu_abi_r2.py tests/test_gpu_ops.py -m gpu -q --timeout 600 2>&1 | tail -3
