#!/bin/bash
# adaptive small side + trimmed builder scan: parity tests, config-3 bench (adaptive / two-step), config 2
O=gpurun_out/r2c5
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -4
for S in 0 1; do
unset SALG_ZSIDE_MODE
if [ "$S" = "1" ]; then export SALG_ZSIDE_MODE=1; fi
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > $O/bench_cfg3_s$S.json 2> $O/bench_cfg3_s$S.err; echo "bench cfg3 variant=$S exit $?"
python - $S <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r2c5/bench_cfg3_s{sys.argv[1]}.json"))
print("ms", round(d["ms_per_step"], 3), "roofline", round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"], 3), "launches", d["gpu_launches"])
print({k: (round(v["ms_total"] / d["steps"], 3), v["launches"] // d["steps"]) for k, v in d["kernel_classes"].items()})
PY
done
unset SALG_ZSIDE_MODE
timeout 900 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-e2e --no-cpu > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench cfg2 exit $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c5/bench_cfg2.json"))
print("cfg2 ms", round(d["ms_per_step"], 3))
print({k: (round(v["ms_total"] / d["steps"], 3), v["launches"] // d["steps"]) for k, v in d["kernel_classes"].items()})
PY
