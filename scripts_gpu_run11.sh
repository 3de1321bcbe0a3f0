#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_pca.py -m gpu -q --timeout 500 2>&1 | tail -5
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-e2e 2>&1 | grep "^{" | python -c "
import json,sys
d=json.loads(sys.stdin.read()); n=d['steps']
print('cfg3 ms_per_step', round(d['ms_per_step'],2)); print({k:(round(v['ms_total']/n,2), v['launches']//n) for k,v in d['kernel_classes'].items()})"
