#!/bin/bash
mkdir -p gpurun_out
SALG_BENCH_VERBOSE=1 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e > gpurun_out/b13a.log 2> gpurun_out/b13a.err; echo "bench exit $?"
grep resident gpurun_out/b13a.err | tail -8
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b13a.log').read().strip().splitlines()[-1])
print(d['ms_per_step'], {k:(round(v['ms_total'],1),v['launches']) for k,v in d['kernel_classes'].items()})
PY
SALG_TC_SORT_BUILD=1 SALG_BENCH_VERBOSE=1 timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu --no-e2e > gpurun_out/b13b.log 2> gpurun_out/b13b.err; echo "bench exit $?"
grep resident gpurun_out/b13b.err | tail -8
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b13b.log').read().strip().splitlines()[-1])
print(d['ms_per_step'], {k:(round(v['ms_total'],1),v['launches']) for k,v in d['kernel_classes'].items()})
PY
