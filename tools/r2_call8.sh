#!/bin/bash
O=gpurun_out/r2c8
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 > $O/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 $O/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "bench cfg3 exit $?"; tail -3 $O/bench_cfg3.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "ref exit $?"; tail -3 $O/bench_ref.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c8/bench_cfg3.json"))
print("cfg3 ms", round(d["ms_per_step"], 2), "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"], 1), "roofline", d["roofline"] and round(d["roofline"]["frac"], 3))
print("cpu", d["cpu_baseline"])
print({k: round(v["ms_total"] / d["steps"], 3) for k, v in d["kernel_classes"].items()})
r = json.load(open("gpurun_out/r2c8/bench_ref.json"))
print("ref", r["ms_per_step"], r["cpu_baseline"]["cores"], r["cpu_baseline"]["phase_seconds_last_step"])
PY
