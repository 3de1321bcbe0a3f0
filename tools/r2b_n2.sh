#!/bin/bash
# 2-GPU check of the peer-memory all-reduce: NCCL parity test of the row-sharded fit, then config 3 with and without it
O=gpurun_out/r2b_n2
mkdir -p $O
echo skip dist
for P in p2p nccl; do
if [ "$P" = "nccl" ]; then export SALG_NO_P2P=1; else unset SALG_NO_P2P; fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu --no-e2e > $O/bench_cfg3_$P.log 2> $O/bench_cfg3_$P.err; echo "bench cfg3 n2 $P exit $?"
python - $P <<'PY'
import json, sys
d=json.loads([l for l in open('gpurun_out/r2b_n2/bench_cfg3_%s.log' % sys.argv[1]) if l.startswith('{')][-1])
n=d['steps']
print(sys.argv[1], 'ms_per_step', round(d['ms_per_step'],3), 'parity', {k: (float('%.2g' % v) if isinstance(v, float) else v) for k, v in (d.get('parity_vs_n1') or {}).items() if k != 'against'})
print({k:(round(v['ms_total']/n,3), v['launches']//n) for k,v in d['kernel_classes'].items() if k in ('allreduce','spmm','spmm_t','gram','stats')})
PY
done
