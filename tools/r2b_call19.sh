#!/bin/bash
O=gpurun_out/r2b19
mkdir -p $O
for S in 0 1; do
if [ "$S" = "1" ]; then export SALG_AX_SINGLE=1; else unset SALG_AX_SINGLE; fi
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e > $O/bench_cfg3_s$S.json 2> $O/bench_cfg3_s$S.err; echo "bench cfg3 single=$S exit $?"
python - $S <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r2b19/bench_cfg3_s{sys.argv[1]}.json"))
print("ms", round(d["ms_per_step"], 2), "roofline", round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"], 3), "parity", d["cpu_baseline"].get("parity_full_operator"))
PY
done
