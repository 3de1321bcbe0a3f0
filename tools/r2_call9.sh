#!/bin/bash
O=gpurun_out/r2c9
mkdir -p $O
for v in o2a o2b o2c o2d; do
SALG_LIB_PATH=scratch/libsalg_$v.so timeout 200 python tools/scripts_tc_diag5.py > $O/$v.log 2>&1; echo "$v $?"
done
cat $O/o2*.log | grep -v "^\[tc" | grep "dbg=  0\|dbg=  2"
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > $O/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -8 $O/pytest_gpu.log
