#!/bin/bash
# 2-GPU validation: partitioned engine vs oracle, replicas bench, partitioned bench (run under gpurun --gpus 2)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
{
timeout 300 $TR --master-port 29511 tests/dist_gpu_check.py 2>&1 | tail -4
timeout 300 $TR --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 5 --no-cpu-baseline 2>/dev/null | tail -1
timeout 300 python bench.py --workload powerlaw10m --scale 0.25 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --workload powerlaw10m --scale 0.25 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1
} > gpurun_out/two_gpu.log 2>&1
cut -c1-700 gpurun_out/two_gpu.log
