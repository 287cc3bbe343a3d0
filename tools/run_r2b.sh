#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_fullsize_configs.py tests/test_gpu_engine.py -q --durations=8 > gpurun_out/r2b_fullsize.log 2>&1; tail -30 gpurun_out/r2b_fullsize.log
