#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_pca.py -m gpu -x -q --timeout 300 2>&1 | tail -2
timeout 300 python scripts_tc_probe.py cfg3 10 2>&1 | tail -2
timeout 300 python scripts_tc_probe.py cfg2 10 2>&1 | tail -2
