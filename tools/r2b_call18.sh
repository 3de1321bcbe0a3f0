#!/bin/bash
O=gpurun_out/r2b18
mkdir -p $O
DBGS=0 timeout 300 python tools/scripts_tm_time.py > $O/plain.log 2>&1 &&
DBGS=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:tm_build -c 2 -o $O/tm_build python tools/scripts_tm_time.py > $O/ncu.log 2>&1
tail -2 $O/ncu.log
