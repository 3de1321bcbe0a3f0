#!/bin/bash
O=gpurun_out/r2c6
mkdir -p $O
for v in g2w8 g2w8b4 g4w4b4 g3w5b4 g2w7 g2w9; do
SALG_LIB_PATH=scratch/libsalg_$v.so timeout 200 python tools/scripts_tc_diag5.py > $O/$v.log 2>&1; echo "$v $?"
done
cat $O/*.log | grep -v "^\[tc" | grep "dbg=  0\|dbg=  2"
