#!/bin/bash
for d in 0 4; do SALG_TC_DBG=$d timeout 300 python scripts_tc_probe.py cfg3 10 2>&1 | tail -2; done
