#!/bin/bash
O=gpurun_out/r2c12
mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_pca.py tests/test_gpu_abi_r2.py -m gpu -q --timeout 600 2>&1 | tail -4
for f in 10 0; do
SALG_JACOBI_DBG=1 SALG_JACOBI_F32=$f timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > $O/b_$f.json 2> $O/b_$f.err
echo "f32=$f: $(tail -1 $O/b_$f.err) $(python -c "
import json; d=json.load(open('$O/b_$f.json')); print(round(d['ms_per_step'],2), {k: round(v['ms_total'] / d['steps'], 3) for k, v in d['kernel_classes'].items()})")"
done
