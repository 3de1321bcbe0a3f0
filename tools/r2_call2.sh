#!/bin/bash
O=gpurun_out/r2c2
mkdir -p $O
timeout 300 python tools/scripts_tc_diag2.py > $O/diag2_default.log 2>&1; echo "diag default $?"
SALG_LIB_PATH=scratch/libsalg_ns7.so timeout 300 python tools/scripts_tc_diag2.py > $O/diag2_ns8.log 2>&1; echo "diag ns8 $?"
SALG_LIB_PATH=scratch/libsalg_ns3.so timeout 300 python tools/scripts_tc_diag2.py > $O/diag2_ns3.log 2>&1; echo "diag ns3 $?"
cat $O/diag2_*.log
