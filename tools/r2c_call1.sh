#!/bin/bash
# fused small-side step: parity tests, then config-3 bench with and without it
O=gpurun_out/r2c1
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_pca.py tests/test_gpu_tm.py tests/test_gpu_ops.py tests/test_gpu_fullsize_parity.py -m gpu -q -x --timeout 600 2>&1 | tail -8
for S in 0 1; do
if [ "$S" = "1" ]; then export SALG_NO_ZSIDE=1; else unset SALG_NO_ZSIDE; fi
timeout 900 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > $O/bench_cfg3_s$S.json 2> $O/bench_cfg3_s$S.err; echo "bench cfg3 nozside=$S exit $?"
python - $S <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r2c1/bench_cfg3_s{sys.argv[1]}.json"))
print("ms", round(d["ms_per_step"], 3), "roofline", round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"], 3), "launches", d["gpu_launches"])
print({k: (round(v["ms_total"] / d["steps"], 3), v["launches"] // d["steps"]) for k, v in d["kernel_classes"].items()})
PY
done
