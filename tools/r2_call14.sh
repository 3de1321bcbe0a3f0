#!/bin/bash
O=gpurun_out/r2c14
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_abi_r2.py -m gpu -q --timeout 600 -x 2>&1 | tail -15
SALG_BENCH_VERBOSE=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; grep "e2e" $O/bench.err | tail -4
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c14/bench.json"))
print("ms", round(d["ms_per_step"], 2), "e2e", d["e2e"])
PY
