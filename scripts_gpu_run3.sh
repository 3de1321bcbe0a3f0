#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q --timeout 900 > gpurun_out/t_dist.log 2>&1; echo "dist exit $?"; tail -30 gpurun_out/t_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_cfg3_n2.log 2>&1; echo "bench n2 exit $?"; tail -c 2500 gpurun_out/bench_cfg3_n2.log
