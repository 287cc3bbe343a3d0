#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2a_gputests.log 2>&1; tail -15 gpurun_out/r2a_gputests.log
timeout 600 python bench.py > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; tail -c 1500 gpurun_out/bench_r2a.json; tail -5 gpurun_out/bench_r2a.err
for cfg in 8x2 16x1@3; do GD_NL_CFG=$cfg timeout 300 python tools/loss_bench.py collab 100 2>&1 | tail -1; done
