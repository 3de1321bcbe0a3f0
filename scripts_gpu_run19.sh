#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -15 > gpurun_out/t_all.log
cat gpurun_out/t_all.log | tail -4
SALG_TC_DBG=32 timeout 300 python scripts_tc_probe.py cfg3 2 2>&1 | grep -v "^$" | tail -4
timeout 300 python scripts_tc_probe.py cfg3 10 2>&1 | tail -2
timeout 300 python scripts_tc_probe.py cfg2 10 2>&1 | tail -2
