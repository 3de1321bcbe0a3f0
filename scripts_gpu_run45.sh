#!/bin/bash
mkdir -p gpurun_out/v3b
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/v3b/bench_cfg3.json 2> gpurun_out/v3b/bench_cfg3.err; echo "bench cfg3 exit $?"
python -c "import json; d=json.loads(open('gpurun_out/v3b/bench_cfg3.json').read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks'])"
