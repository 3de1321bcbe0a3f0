#!/bin/bash
mkdir -p gpurun_out
for dbg in 32 423; do SALG_TC_DBG=$dbg timeout 300 python scripts_tc_probe.py cfg3 1 2>&1 | grep -v Warn | grep "ax\]\|False" | tail -3; done | tee gpurun_out/tc_probe.log
