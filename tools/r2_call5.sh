#!/bin/bash
O=gpurun_out/r2c5
mkdir -p $O
timeout 200 python tools/scripts_tc_diag5.py > $O/default.log 2>&1; echo "default $?"
for v in g1w15 g1w15b2 g1w16 g2w8 g1w10; do
SALG_LIB_PATH=scratch/libsalg_$v.so timeout 200 python tools/scripts_tc_diag5.py > $O/$v.log 2>&1; echo "$v $?"
done
cat $O/*.log | grep -v "^\[tc"
