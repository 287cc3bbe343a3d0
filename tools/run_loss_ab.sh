#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/loss_ab.log
timeout 300 python -m pytest tests/test_gpu_gcn.py -q -k "fused_edge_loss" >> gpurun_out/loss_ab.log 2>&1
for cfg in 8x2 16x1@4 16x1@3; do
  GD_NL_CFG=$cfg timeout 300 python tools/loss_bench.py collab 100 2>&1 | tail -1 >> gpurun_out/loss_ab.log
done
cat gpurun_out/loss_ab.log | tail -8
