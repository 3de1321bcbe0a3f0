#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts_chol_probe.py > gpurun_out/chol_plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:chol_inv -s 2 -c 1 -o gpurun_out/prof_chol python scripts_chol_probe.py > gpurun_out/ncu_chol.log 2>&1; echo "ncu exit $?"
