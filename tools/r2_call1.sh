#!/bin/bash
O=gpurun_out/r2c1
mkdir -p $O
(nproc; free -g; lscpu | head -30; nvidia-smi; nvidia-smi topo -m) > $O/host.log 2>&1
timeout 300 python tools/scripts_tc_diag.py > $O/diag_default.log 2>&1; echo "diag default $?"
SALG_TC_ROT=1 timeout 300 python tools/scripts_tc_diag.py > $O/diag_rot.log 2>&1; echo "diag rot $?"
SALG_LIB_PATH=scratch/libsalg_unsc.so SALG_TC_ROT=1 timeout 300 python tools/scripts_tc_diag.py > $O/diag_unsc.log 2>&1; echo "diag unsc $?"
cat $O/diag_*.log
