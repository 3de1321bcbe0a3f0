#!/bin/bash
# Round-1 v4 measurement set (one B200): tests, bench cfg3 / cfg2 / reference arm, ncu launch list, secondary configs,
# the reference's own PCA test at full size through the C++ facade.
# Every ncu pass follows a plain run of the same command that exited 0.
O=gpurun_out/v4f
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 $O/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "bench cfg3 exit $?"
timeout 600 python bench.py --workload cfg2 --steps 10 --warmup 3 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench cfg2 exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_ref.err; echo "ref exit $?"
timeout 600 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $O/plain_launch.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_cfg3.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $O/ncu_launch.log 2>&1; echo "ncu launches exit $?"
timeout 600 python bench_extra.py > $O/extra.log 2>&1; echo "extra exit $?"
( time timeout 600 ./tests/cpp/build/facade_test --full ) > $O/facade_full.log 2>&1; echo "facade full exit $?"
python - <<'PY'
import json
for f in ("bench_cfg3", "bench_cfg2"):
    d = json.load(open(f"gpurun_out/v4f/{f}.json"))
    print(f, round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 1), "roofline", round(d["roofline"]["frac"], 3), d["e2e"]["class_ms_per_step"])
PY
cut -c1-200 $O/extra.log
