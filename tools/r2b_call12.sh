#!/bin/bash
# state check of the TMEM-operand products: parity tests, benchmark lines (config 3, config 2), launch list, ncu --set full
O=gpurun_out/r2b12
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_tm.py tests/test_gpu_ops.py tests/test_gpu_fullsize_parity.py -m gpu -q -x --timeout 600 2>&1 | tail -4
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "bench cfg3 exit $?"
timeout 900 python bench.py --workload cfg2 --steps 5 --warmup 3 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench cfg2 exit $?"
python - <<'PY'
import json
for w in ("cfg3", "cfg2"):
    try:
        d = json.load(open(f"gpurun_out/r2b12/bench_{w}.json"))
        print(w, "ms", round(d["ms_per_step"], 2), "e2e ms", round(d["e2e"]["ms_per_step"], 1) if d.get("e2e") else None, "roofline", round(d["roofline"]["frac"], 3), d["roofline"]["kernel"][:12], round(d["roofline"]["avg_launch_ms"], 3))
        print({k: (round(v["ms_total"] / d["steps"], 3), v["launches"] // d["steps"], round(v["frac_of_hbm_peak"], 3) if v["frac_of_hbm_peak"] else None) for k, v in d["kernel_classes"].items()})
    except Exception as e:
        print(w, "failed", e)
PY
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity > $O/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_cfg3.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity > $O/ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tm_product|tm_build" -s 4 -c 6 -o $O/tm_kernels python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity > $O/ncu2.log 2>&1
tail -2 $O/ncu2.log
