#!/bin/bash
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -3
for i in 1 2; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('resident', round(d['ms_per_step'],2), d['gpu_launches'], round(sum(v['ms_total'] for v in d['kernel_classes'].values())/d['steps'],2))"
done
timeout 600 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg2 resident', round(d['ms_per_step'],2))"
