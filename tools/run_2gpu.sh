#!/bin/bash
# 2-GPU validation of the row-partitioned epoch: NCCL tests, then the bench line with a reduced config-5 graph.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dist.py -q > gpurun_out/two_gpu_tests.log 2>&1; tail -6 gpurun_out/two_gpu_tests.log
SCALE=${1:-0.1}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 \
    bench.py --gpus 2 --steps 20 --warmup 5 --partition-scale $SCALE > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err
tail -c 3000 gpurun_out/bench_2gpu.json; tail -8 gpurun_out/bench_2gpu.err
