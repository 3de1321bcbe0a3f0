#!/bin/bash
# final measurement set of round 2, third session (one B200): whole GPU suite, benchmark lines, launch list, ncu --set full
O=gpurun_out/r2c_final
mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -6
timeout 900 python bench.py --steps 5 --warmup 3 > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "bench cfg3 exit $?"
timeout 900 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench cfg2 exit $?"
timeout 900 python bench.py --workload cfg5 --steps 3 --warmup 3 --no-cpu > $O/bench_cfg5.json 2> $O/bench_cfg5.err; echo "bench cfg5 exit $?"
python - <<'PY'
import json
for w in ("cfg3", "cfg2", "cfg5"):
    try:
        d = json.load(open(f"gpurun_out/r2c_final/bench_{w}.json"))
        print(w, "ms", round(d["ms_per_step"], 2), "e2e ms", round(d["e2e"]["ms_per_step"], 1) if d.get("e2e") else None, "roofline", round(d["roofline"]["frac"], 3), round(d["roofline"]["avg_launch_ms"], 3), "cpu", d.get("cpu_baseline") and d["cpu_baseline"].get("value"), "launches", d["gpu_launches"])
        print({k: (round(v["ms_total"] / d["steps"], 3), v["launches"] // d["steps"], round(v["frac_of_hbm_peak"], 3) if v["frac_of_hbm_peak"] else None) for k, v in d["kernel_classes"].items()})
    except Exception as e:
        print(w, "failed", e)
PY
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity > $O/plain.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_cfg3.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity > $O/ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tm_product|tm_build2|zside_solve|tm_zside_apply|col_stats_masked|tc_gram_prep|jacobi_svd64" -c 10 -o $O/top_kernels python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-parity > $O/ncu2.log 2>&1
tail -2 $O/ncu2.log
timeout 300 ncu -i $O/top_kernels.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,lts__t_sectors.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active > $O/top_kernels_raw.csv 2>/dev/null
wc -l $O/top_kernels_raw.csv
