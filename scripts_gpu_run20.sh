#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -15 > gpurun_out/t_all.log
cat gpurun_out/t_all.log | tail -4
timeout 600 python bench.py --steps 6 --warmup 3 --no-e2e > gpurun_out/b20.log 2> gpurun_out/b20.err; echo "bench exit $?"
tail -3 gpurun_out/b20.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b20.log').read().strip().splitlines()[-1])
print(d['ms_per_step'], {k:(round(v['ms_total']/d['steps'],2),v['launches']//d['steps']) for k,v in d['kernel_classes'].items()})
print(d['cpu_baseline']['parity_on_sample'])
PY
