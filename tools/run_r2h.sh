#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gpu_engine.py -x -q -m gpu -k "dense" > gpurun_out/r2h_tests.log 2>&1; tail -15 gpurun_out/r2h_tests.log
timeout 60 python tools/dense_ni_bench.py 2>&1 | tee gpurun_out/r2h_dense_ni.log | tail -3
timeout 90 python -m pytest tests/test_gpu_fullsize_configs.py -x -q -m gpu -k "dense" > gpurun_out/r2h_full.log 2>&1; tail -5 gpurun_out/r2h_full.log
