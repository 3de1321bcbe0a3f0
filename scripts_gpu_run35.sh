#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -x -q --timeout 600 -k "config4 or config5" 2>&1 | tail -25
