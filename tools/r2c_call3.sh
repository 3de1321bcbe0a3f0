#!/bin/bash
# two-step fused small side (default) + dense-image tile builder: parity tests, config-3 bench (new / first builder), cfg2
O=gpurun_out/r2c3
mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_tm.py tests/test_gpu_ops.py tests/test_gpu_pca.py tests/test_gpu_fullsize_parity.py tests/test_gpu_scale.py -m gpu -q -x --timeout 600 2>&1 | tail -12
for S in 0 1; do
unset SALG_TM_BUILD
if [ "$S" = "1" ]; then export SALG_TM_BUILD=1; fi
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > $O/bench_cfg3_s$S.json 2> $O/bench_cfg3_s$S.err; echo "bench cfg3 variant=$S exit $?"
python - $S <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r2c3/bench_cfg3_s{sys.argv[1]}.json"))
print("ms", round(d["ms_per_step"], 3), "roofline", round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"], 3), "launches", d["gpu_launches"])
print({k: (round(v["ms_total"] / d["steps"], 3), v["launches"] // d["steps"]) for k, v in d["kernel_classes"].items()})
PY
done
unset SALG_TM_BUILD
timeout 900 python bench.py --workload cfg5 --steps 2 --warmup 2 --no-e2e --no-cpu > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "bench cfg5 exit $?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2c3/bench_cfg5.json"))
print("cfg5 ms", round(d["ms_per_step"], 3))
print({k: (round(v["ms_total"] / d["steps"], 3), v["launches"] // d["steps"]) for k, v in d["kernel_classes"].items()})
PY
