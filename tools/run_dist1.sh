#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_dist.py -q -x > gpurun_out/dist1.log 2>&1; tail -25 gpurun_out/dist1.log
ncu --set full --clock-control none --import-source on -k regex:node_loss_kernel -s 3 -c 1 -f -o gpurun_out/r2_node_loss_v2 \
    python tools/loss_bench.py collab 3 > gpurun_out/r2_node_loss_v2.log 2>&1
tail -2 gpurun_out/r2_node_loss_v2.log
