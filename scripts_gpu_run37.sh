#!/bin/bash
python scripts_tc_probe2.py 2>&1 | tail -1
SALG_LIB_PATH=scratch/libsalg_fw.so python scripts_tc_probe2.py 2>&1 | tail -1
