#!/bin/bash
O=gpurun_out/r2c10
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_pca.py tests/test_gpu_abi_r2.py -m gpu -q --timeout 600 > $O/pytest.log 2>&1; echo "pytest exit $?"; tail -6 $O/pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > $O/bench.json 2> $O/bench.err; echo "bench exit $?"; tail -3 $O/bench.err
SALG_JACOBI_F32=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu > $O/bench_f64jac.json 2> $O/bench2.err; echo "bench2 exit $?"
python - <<'PY'
import json
for f in ("bench", "bench_f64jac"):
    d = json.load(open(f"gpurun_out/r2c10/{f}.json"))
    print(f, "ms", round(d["ms_per_step"], 2), {k: round(v["ms_total"] / d["steps"], 3) for k, v in d["kernel_classes"].items()})
PY
