#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -15 > gpurun_out/t_all.log
cat gpurun_out/t_all.log | tail -4
timeout 300 python scripts_tc_probe.py cfg3 10 2>&1 | tail -2
timeout 300 python scripts_tc_probe.py cfg2 3 2>&1 | tail -2
timeout 600 python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu > gpurun_out/b23.log 2> gpurun_out/b23.err; echo "bench exit $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/b23.log').read().strip().splitlines()[-1])
print(d['ms_per_step'], {k:(round(v['ms_total']/d['steps'],2),v['launches']//d['steps']) for k,v in d['kernel_classes'].items()})
PY
