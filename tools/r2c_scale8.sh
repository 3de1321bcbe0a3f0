#!/bin/bash
# 8-GPU run of the final code: config 3 (strong scaling, e2e, parity of the 8-rank fit against a single-GPU fit in the line)
N=8
O=gpurun_out/r2c_n8
mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu > $O/bench_cfg3.log 2> $O/bench_cfg3.err; echo "bench cfg3 n$N exit $?"; tail -2 $O/bench_cfg3.err
python - <<PY
import json
for w in ("cfg3",):
    try:
        d=json.loads([l for l in open('$O/bench_%s.log' % w) if l.startswith('{')][-1])
    except Exception as e:
        print(w, "no line", e); continue
    print(w, 'n', d['n_gpus'], 'ms_per_step', round(d['ms_per_step'],2), 'value', round(d['value']), 'e2e', d['e2e'] and d['e2e'].get('ms_per_step') and round(d['e2e']['ms_per_step'],1), 'parity', d.get('parity_vs_n1'), 'launches', d['gpu_launches'])
    n=d['steps']
    print({k:(round(v['ms_total']/n,2), v['launches']//n, v.get('frac_of_hbm_peak') and round(v['frac_of_hbm_peak'],3)) for k,v in d['kernel_classes'].items()})
PY
