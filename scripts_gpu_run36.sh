#!/bin/bash
python scripts_tc_probe2.py 2>&1 | tail -1
SALG_TC_ORDER=1 python scripts_tc_probe2.py 2>&1 | tail -1
for v in u1w5 u1w4 u0w4; do
SALG_LIB_PATH=scratch/libsalg_$v.so python scripts_tc_probe2.py 2>&1 | tail -1
SALG_TC_ORDER=1 SALG_LIB_PATH=scratch/libsalg_$v.so python scripts_tc_probe2.py 2>&1 | tail -1
done
