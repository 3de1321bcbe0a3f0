#!/bin/bash
SALG_TC_DBG=32 timeout 300 python scripts_tc_probe.py cfg3 3 2>&1 | grep -v "^$" | tail -8
