#!/bin/bash
for i in 1 2 3; do
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('with clocks', round(d['ms_per_step'],2), d['clocks']['samples'], round(sum(v['ms_total'] for v in d['kernel_classes'].values())/d['steps'],2))"
SALG_BENCH_NO_CLOCKS=1 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('no clocks  ', round(d['ms_per_step'],2), d['clocks']['samples'], round(sum(v['ms_total'] for v in d['kernel_classes'].values())/d['steps'],2))"
done
