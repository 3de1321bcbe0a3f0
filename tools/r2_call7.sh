#!/bin/bash
O=gpurun_out/r2c7
mkdir -p $O
M=gpu__time_duration.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__inst_executed.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed
export SALG_LIB_PATH=scratch/libsalg_g3w5b4.so
python tools/scripts_tc_one.py > $O/plain0.log 2>&1 && ncu --metrics $M --clock-control none -k regex:"tc_ax_kernel|tc_aty_kernel" -s 6 -c 2 --csv --log-file $O/ncu_rot0.csv python tools/scripts_tc_one.py > $O/ncu0.log 2>&1; echo "rot0 $?"
export SALG_TC_ROT=1
python tools/scripts_tc_one.py > $O/plain1.log 2>&1 && ncu --metrics $M --clock-control none -k regex:"tc_ax_kernel|tc_aty_kernel" -s 6 -c 2 --csv --log-file $O/ncu_rot1.csv python tools/scripts_tc_one.py > $O/ncu1.log 2>&1; echo "rot1 $?"
unset SALG_TC_ROT; export SALG_TC_ORDER=1
python tools/scripts_tc_one.py > $O/plain2.log 2>&1 && ncu --metrics $M --clock-control none -k regex:"tc_ax_kernel|tc_aty_kernel" -s 6 -c 2 --csv --log-file $O/ncu_ord.csv python tools/scripts_tc_one.py > $O/ncu2.log 2>&1; echo "order $?"
cat $O/plain*.log; tail -n 20 $O/ncu_rot0.csv $O/ncu_rot1.csv $O/ncu_ord.csv | cut -c1-300
