#!/bin/bash
# last check of the round: smoke(), the whole GPU suite, one default bench line
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg3 ms', round(d['ms_per_step'],3), 'launches', d['gpu_launches'], 'roofline', round(d['roofline']['frac'],3))"
