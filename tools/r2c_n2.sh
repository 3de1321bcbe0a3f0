#!/bin/bash
# 2-GPU: NCCL parity test of the row-sharded fit (fused small side), then the config-3 line
O=gpurun_out/r2c_n2
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q -s --timeout 600 2>&1 | tail -12 | tee $O/dist_gpu_check.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-e2e > $O/bench_cfg3.log 2> $O/bench_cfg3.err; echo "bench cfg3 n2 exit $?"
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2c_n2/bench_cfg3.log') if l.startswith('{')][-1])
n=d['steps']
print('ms_per_step', round(d['ms_per_step'],3), 'parity', {k: (float('%.2g' % v) if isinstance(v, float) else v) for k, v in (d.get('parity_vs_n1') or {}).items() if k != 'against'})
print({k:(round(v['ms_total']/n,3), v['launches']//n) for k,v in d['kernel_classes'].items()})
PY
