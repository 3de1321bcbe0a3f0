#!/bin/bash
O=gpurun_out/r2b16
mkdir -p $O
for W in cfg5 cfg2; do
SALG_SPMM_IMPL=tc timeout 1200 python bench.py --workload $W --steps 3 --warmup 3 --no-cpu --no-e2e > $O/bench_${W}_tc.json 2> $O/bench_${W}_tc.err; echo "bench $W tc exit $?"
python - $W <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r2b16/bench_{sys.argv[1]}_tc.json"))
print("ms", round(d["ms_per_step"], 2))
print({k: (round(v["ms_total"] / d["steps"], 3), v["launches"] // d["steps"], round(v["frac_of_hbm_peak"], 3) if v["frac_of_hbm_peak"] else None) for k, v in d["kernel_classes"].items()})
PY
done
