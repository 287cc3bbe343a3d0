#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_gcn.py -x -q -m gpu -k "gemm" > gpurun_out/r2n_tests.log 2>&1; tail -12 gpurun_out/r2n_tests.log
timeout 120 python tools/gemm_ab.py 2>&1 | tee gpurun_out/r2n_gemm_ab.log | tail -8
