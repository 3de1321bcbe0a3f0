#!/bin/bash
python scripts_stats_probe.py 2>&1 | tail -1
for v in 1 2 3; do SALG_LIB_PATH=scratch/libsalg_s$v.so python scripts_stats_probe.py 2>&1 | tail -1; done
