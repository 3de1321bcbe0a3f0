#!/bin/bash
mkdir -p gpurun_out
for f in csr stats ops pca scale; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -m gpu -q --timeout 600 > gpurun_out/t_$f.log 2>&1
  echo "$f exit $?"; tail -3 gpurun_out/t_$f.log
done
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_cfg3.log 2>gpurun_out/bench_cfg3.err; echo "bench cfg3 exit $?"
timeout 900 python bench.py --workload cfg2 --steps 5 --warmup 3 > gpurun_out/bench_cfg2.log 2>&1; echo "bench cfg2 exit $?"
python - <<'PY'
import json
for f in ('gpurun_out/bench_cfg3.log','gpurun_out/bench_cfg2.log'):
    d=json.loads([l for l in open(f) if l.startswith('{')][-1])
    print(f, 'ms_per_step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],1), 'roofline', round(d['roofline']['frac'],3), d['roofline']['kernel'][:20], 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value']))
    n=d['steps']
    print({k:(round(v['ms_total']/n,2), v['launches']//n) for k,v in d['kernel_classes'].items()})
PY
