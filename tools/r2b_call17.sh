#!/bin/bash
O=gpurun_out/r2b17
mkdir -p $O
for C in c8 c16; do
SALG_LIB_PATH=scratch/libsalg_$C.so timeout 1200 python bench.py --workload cfg5 --steps 2 --warmup 2 --no-cpu --no-e2e > $O/bench_cfg5_$C.json 2> $O/bench_cfg5_$C.err; echo "bench cfg5 $C exit $?"
python - $C <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/r2b17/bench_cfg5_{sys.argv[1]}.json"))
print(sys.argv[1], "ms", round(d["ms_per_step"], 2), {k: (round(v["ms_total"] / d["steps"], 3)) for k, v in d["kernel_classes"].items() if k in ("transpose",)})
PY
done
