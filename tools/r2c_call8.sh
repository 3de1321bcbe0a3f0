#!/bin/bash
O=gpurun_out/r2c8
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -25
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "bench cfg3 exit $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c8/bench_cfg3.json"))
print("ms", round(d["ms_per_step"], 3), "roofline", round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"], 3), "launches", d["gpu_launches"])
print({k: (round(v["ms_total"] / d["steps"], 3), v["launches"] // d["steps"]) for k, v in d["kernel_classes"].items()})
PY
