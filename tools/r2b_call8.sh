#!/bin/bash
# whole GPU suite with the TMEM-operand products as the default, then the benchmark lines
O=gpurun_out/r2b8
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -6
SALG_BENCH_VERBOSE=1 timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2b8/bench_cfg3.json"))
print("ms", round(d["ms_per_step"], 2), "e2e", d["e2e"], "roofline", d["roofline"], "cpu", d.get("cpu_baseline", {}).get("value"))
print({k: d[k] for k in d if k not in ("roofline", "e2e", "cpu_baseline", "config")})
PY
