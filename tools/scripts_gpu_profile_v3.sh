#!/bin/bash
# Round-1 v3 measurement set (one B200): tests, bench cfg3 / cfg2 / reference arm, ncu launch list, ncu --set full of the
# dominant kernels.  Every ncu pass follows a plain run of the same command that exited 0.
mkdir -p gpurun_out/v3b
O=gpurun_out/v3b
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > $O/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 $O/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $O/bench_cfg3.json 2> $O/bench_cfg3.err; echo "bench cfg3 exit $?"
timeout 600 python bench.py --workload cfg2 --steps 10 --warmup 3 > $O/bench_cfg2.json 2> $O/bench_cfg2.err; echo "bench cfg2 exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_ref.err; echo "ref exit $?"
timeout 600 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $O/plain_launch.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $O/launches_cfg3.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu > $O/ncu_launch.log 2>&1; echo "ncu launches exit $?"
timeout 300 python tools/scripts_tc_probe.py cfg3 1 > $O/probe_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_ax_kernel|tc_aty_kernel" -s 3 -c 2 -o $O/prof_tc -f python tools/scripts_tc_probe.py cfg3 1 > $O/ncu_tc.log 2>&1; echo "ncu tc exit $?"
timeout 900 ncu --set full --clock-control none -k regex:"col_stats_masked|tc_bin_kernel" -s 2 -c 2 -o $O/prof_misc -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $O/ncu_misc.log 2>&1; echo "ncu misc exit $?"
python bench_extra.py > $O/extra.log 2>&1; echo "extra exit $?"
