#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout 200 -x > gpurun_out/t_ops.log 2>&1; echo "ops exit $?"; tail -25 gpurun_out/t_ops.log
