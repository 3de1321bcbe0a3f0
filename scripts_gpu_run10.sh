#!/bin/bash
mkdir -p gpurun_out
for f in csr stats ops pca scale; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -m gpu -q --timeout 600 > gpurun_out/t_$f.log 2>&1
  echo "$f exit $?"; tail -2 gpurun_out/t_$f.log
done
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_cfg3.log 2>gpurun_out/bench_cfg3.err; echo "bench cfg3 exit $?"
timeout 900 python bench.py --workload cfg2 --steps 10 --warmup 3 > gpurun_out/bench_cfg2.log 2>&1; echo "bench cfg2 exit $?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "bench ref exit $?"
# launch list of the bench command (plain run first)
timeout 900 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/plain.log 2>&1 && \
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_cfg3.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu.log 2>&1; echo "ncu list exit $?"
# full capture of the dominant kernels (one launch each), plain run first
timeout 300 python scripts_tc_probe.py cfg3 1 > gpurun_out/probe_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_a -s 4 -c 2 -o gpurun_out/prof_tc python scripts_tc_probe.py cfg3 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu full exit $?"
python - <<'PY'
import json
for f in ('gpurun_out/bench_cfg3.log','gpurun_out/bench_cfg2.log'):
    d=json.loads([l for l in open(f) if l.startswith('{')][-1])
    print(f, 'ms_per_step', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['ms_per_step'],1), 'roofline', round(d['roofline']['frac'],3), d['roofline']['kernel'][:20], 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value']))
    n=d['steps']
    print({k:(round(v['ms_total']/n,2), v['launches']//n) for k,v in d['kernel_classes'].items()})
PY
