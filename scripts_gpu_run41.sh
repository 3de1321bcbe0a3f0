#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -5
timeout 600 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu > gpurun_out/b41.log 2> gpurun_out/b41.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b41.log').read().strip().splitlines()[-1])
print(d['ms_per_step'], {k:(round(v['ms_total']/d['steps'],2),v['launches']//d['steps']) for k,v in d['kernel_classes'].items()})
PY
