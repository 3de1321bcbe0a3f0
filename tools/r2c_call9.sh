#!/bin/bash
# adaptive small side v2 (two steps first, one step attempted afterwards) against the two-step default
O=gpurun_out/r2c9
mkdir -p $O
SALG_ZSIDE_MODE=0 timeout 1500 python -m pytest tests/test_gpu_fused_side.py tests/test_gpu_pca.py tests/test_gpu_scale.py tests/test_gpu_fullsize_parity.py -m gpu -q --timeout 900 2>&1 | tail -4
for W in cfg3 cfg2; do
for S in 1 0; do
export SALG_ZSIDE_MODE=$S
timeout 900 python bench.py --workload $W --steps 5 --warmup 3 --no-e2e --no-cpu > $O/bench_${W}_s$S.json 2> $O/bench_${W}_s$S.err; echo "bench $W mode=$S exit $?"
python - $W $S <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r2c9/bench_{sys.argv[1]}_s{sys.argv[2]}.json"))
print("ms", round(d["ms_per_step"], 3), "launches", d["gpu_launches"], {k: (round(v["ms_total"] / d["steps"], 3), v["launches"] // d["steps"]) for k, v in d["kernel_classes"].items() if k in ("chol", "gram", "panel_mul")})
PY
done
done
