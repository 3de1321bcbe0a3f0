#!/bin/bash
mkdir -p gpurun_out
echo "== no fused"; SALG_NO_FUSED_NORM=1 timeout 300 python -m pytest tests/test_gpu_pca.py -m gpu -x -q --timeout 200 2>&1 | tail -3
echo "== fused"; timeout 300 python -m pytest tests/test_gpu_pca.py -m gpu -x -q --timeout 200 -k test_f32_randomized_against_f64_oracle 2>&1 | tail -3
echo "== sanitizer"; timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_pca.py -m gpu -x -q --timeout 500 -k test_f32_randomized_against_f64_oracle 2>&1 | grep -v "^$" | head -40
