#!/bin/bash
O=gpurun_out/r2b15
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_tm.py tests/test_gpu_pca.py -m gpu -q -x --timeout 300 2>&1 | tail -3
for W in cfg3; do
timeout 1200 python bench.py --workload $W --steps 5 --warmup 3 --no-cpu --no-e2e > $O/bench_${W}_n1.json 2> $O/bench_${W}_n1.err; echo "bench $W exit $?"
python - $W <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r2b15/bench_{sys.argv[1]}_n1.json"))
print("ms", round(d["ms_per_step"], 2), "roofline", d["roofline"]["kernel"][:14], round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"], 3))
print({k: (round(v["ms_total"] / d["steps"], 3), v["launches"] // d["steps"], round(v["frac_of_hbm_peak"], 3) if v["frac_of_hbm_peak"] else None) for k, v in d["kernel_classes"].items()})
PY
done
