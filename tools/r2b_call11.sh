#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_tm.py -m gpu -q -x --timeout 300 2>&1 | grep -E "Error|assert|error|passed|failed" | head -20
