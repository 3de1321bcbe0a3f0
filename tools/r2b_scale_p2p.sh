#!/bin/bash
# N-GPU run: config 3 with the peer-memory all-reduce and with NCCL
N=$1
O=gpurun_out/r2b_p2p_n$N
mkdir -p $O
for P in p2p nccl; do
if [ "$P" = "nccl" ]; then export SALG_P2P=0; else export SALG_P2P=1; fi
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu --no-e2e > $O/bench_cfg3_$P.log 2> $O/bench_cfg3_$P.err; echo "bench cfg3 n$N $P exit $?"
python - $P $N <<'PY'
import json, sys
d=json.loads([l for l in open('gpurun_out/r2b_p2p_n%s/bench_cfg3_%s.log' % (sys.argv[2], sys.argv[1])) if l.startswith('{')][-1])
n=d['steps']
print(sys.argv[1], 'n', d['n_gpus'], 'ms_per_step', round(d['ms_per_step'],3), 'parity', {k: (float('%.2g' % v) if isinstance(v, float) else v) for k, v in (d.get('parity_vs_n1') or {}).items() if k != 'against'})
print({k:(round(v['ms_total']/n,3), v['launches']//n) for k,v in d['kernel_classes'].items()})
PY
done
