#!/bin/bash
# final measurement set of round 2 (one B200): whole GPU suite, benchmark lines, launch list, ncu --set full of the top kernels
O=gpurun_out/r2b_final
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -5
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "bench cfg3 exit $?"
timeout 900 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench cfg2 exit $?"
timeout 900 python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "bench cfg5 exit $?"
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > $O/bench_reference.json 2> $O/bench_reference.err; echo "bench reference exit $?"
python - <<'PY'
import json
for w in ("cfg3", "cfg2", "cfg5"):
    try:
        d = json.load(open(f"gpurun_out/r2b_final/bench_{w}.json"))
        print(w, "ms", round(d["ms_per_step"], 2), "e2e ms", round(d["e2e"]["ms_per_step"], 1) if d.get("e2e") else None, "roofline", round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"], 3), "cpu", d.get("cpu_baseline") and d["cpu_baseline"].get("value"))
        print({k: (round(v["ms_total"] / d["steps"], 3), v["launches"] // d["steps"], round(v["frac_of_hbm_peak"], 3) if v["frac_of_hbm_peak"] else None) for k, v in d["kernel_classes"].items()})
    except Exception as e:
        print(w, "failed", e)
try:
    d = json.load(open("gpurun_out/r2b_final/bench_reference.json")); print("reference", d.get("value"), d.get("unit"), d.get("ms_per_step"), d.get("cpu_baseline"))
except Exception as e:
    print("reference failed", e)
PY
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity > $O/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_cfg3.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity > $O/ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tm_product|tm_build|col_stats_masked|tc_gram_prep" -c 9 -o $O/top_kernels python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity > $O/ncu2.log 2>&1
tail -2 $O/ncu2.log
