#!/bin/bash
python scripts_stats_probe.py 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q --timeout 600 2>&1 | tail -4
