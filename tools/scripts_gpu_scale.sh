#!/bin/bash
# usage: scripts_gpu_scale.sh N
N=$1
mkdir -p gpurun_out
if [ "$N" = "2" ]; then
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 900 > gpurun_out/t_dist.log 2>&1; echo "dist exit $?"; tail -3 gpurun_out/t_dist.log
fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_cfg3_n$N.log 2>&1; echo "bench n$N exit $?"
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_cfg3_n$N.log') if l.startswith('{')][-1])
print('n', d['n_gpus'], 'ms_per_step', round(d['ms_per_step'],2), 'e2e', d['e2e'] and round(d['e2e']['ms_per_step'],1))
n=d['steps']
print({k:(round(v['ms_total']/n,2), v['launches']//n) for k,v in d['kernel_classes'].items()})
PY
