#!/bin/bash
mkdir -p gpurun_out
for f in csr stats pca; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -m gpu -q --timeout 600 > gpurun_out/t_$f.log 2>&1
  echo "$f exit $?"; tail -4 gpurun_out/t_$f.log
done
SALG_BENCH_VERBOSE=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/bench_cfg3.log 2> gpurun_out/bench_cfg3.err; echo "bench cfg3 exit $?"; grep e2e gpurun_out/bench_cfg3.err | tail -4; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_cfg3.log') if l.startswith('{')][-1])
print('ms_per_step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])
print({k:(round(v['ms_total']/3,2), v['launches']//3) for k,v in d['kernel_classes'].items()})
print(d['e2e'].get('class_ms_per_step'))
PY
