#!/bin/bash
# ncu --set full of the kernels rewritten in v4 (each capture follows a plain run of the same command that exited 0).
O=gpurun_out/v4n; mkdir -p $O
timeout 200 python bench_extra.py cfg5 > $O/extra_plain.log 2>&1 && \
timeout 400 ncu --set full --clock-control none -k regex:"sum_row_kernel|log1p_kernel|normalize_row_kernel|preprocess_row_kernel|col_stats_tiled_kernel" -c 16 -o $O/prof_stream -f python bench_extra.py cfg5 > $O/ncu_stream.log 2>&1; echo "ncu stream exit $?"
ncu -i $O/prof_stream.ncu-rep --page raw --csv > $O/prof_stream_raw.csv 2>/dev/null
timeout 200 python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $O/plain.log 2>&1 && \
timeout 400 ncu --set full --clock-control none -k regex:"col_stats_masked|tc_bin_kernel" -s 2 -c 2 -o $O/prof_misc -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > $O/ncu_misc.log 2>&1; echo "ncu misc exit $?"
ncu -i $O/prof_misc.ncu-rep --page raw --csv > $O/prof_misc_raw.csv 2>/dev/null
ls -la $O
