#!/bin/bash
# Jacobi early exit + tail overlap: parity tests (all -m gpu), config-3 bench variants, sweep counts
O=gpurun_out/r2c4
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -8
for S in 0 1 2; do
unset SALG_NO_TAIL_OVERLAP SALG_JACOBI_CONFIRM
if [ "$S" = "1" ]; then export SALG_NO_TAIL_OVERLAP=1; fi
if [ "$S" = "2" ]; then export SALG_NO_TAIL_OVERLAP=1 SALG_JACOBI_CONFIRM=1; fi
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > $O/bench_cfg3_s$S.json 2> $O/bench_cfg3_s$S.err; echo "bench cfg3 variant=$S exit $?"
python - $S <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r2c4/bench_cfg3_s{sys.argv[1]}.json"))
print("ms", round(d["ms_per_step"], 3), "roofline", round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"], 3), "launches", d["gpu_launches"], "parity", d.get("parity"))
print({k: (round(v["ms_total"] / d["steps"], 3), v["launches"] // d["steps"]) for k, v in d["kernel_classes"].items()})
PY
done
unset SALG_NO_TAIL_OVERLAP SALG_JACOBI_CONFIRM
SALG_JACOBI_DBG=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-parity 2>&1 | grep jacobi | head -2
