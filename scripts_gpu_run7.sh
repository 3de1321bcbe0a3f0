#!/bin/bash
mkdir -p gpurun_out
for f in csr stats ops pca; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -m gpu -q --timeout 600 > gpurun_out/t_$f.log 2>&1
  echo "$f exit $?"; tail -2 gpurun_out/t_$f.log
done
SALG_BENCH_VERBOSE=1 timeout 900 python bench.py --steps 6 --warmup 3 --no-cpu > gpurun_out/bench_cfg3.log 2> gpurun_out/bench_cfg3.err; echo "bench exit $?"; grep "resident\|e2e" gpurun_out/bench_cfg3.err | tail -12
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/bench_cfg3.log') if l.startswith('{')][-1])
print('ms_per_step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],1))
PY
