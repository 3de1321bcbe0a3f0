#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -12
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_cfg3_n2b.log 2>&1; echo "bench n2 exit $?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_cfg3_n2b.log') if l.startswith('{')][-1])
n=d['steps']
print('n', d['n_gpus'], 'ms_per_step', round(d['ms_per_step'],2), {k:(round(v['ms_total']/n,2), v['launches']//n) for k,v in d['kernel_classes'].items()})
PY
