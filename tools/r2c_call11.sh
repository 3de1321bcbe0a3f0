#!/bin/bash
O=gpurun_out/r2c11
mkdir -p $O
timeout 100 ncu --set full --clock-control none -k regex:spmm_chunk -c 4 -o $O/f64_spmm python tools/probe_f64_spmm_ncu.py > $O/ncu.log 2>&1
tail -3 $O/ncu.log
timeout 60 ncu -i $O/f64_spmm.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,l1tex__t_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sectors.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,launch__grid_size,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active > $O/f64_spmm_raw.csv 2>/dev/null
cat $O/f64_spmm_raw.csv | cut -c1-600 | tail -5
