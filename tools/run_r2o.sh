#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gcn.py tests/test_gpu_engine.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -3
GD_LIB_TAG=new timeout 120 python tools/epoch_ab.py 2>&1 | tail -1
timeout 120 python tools/gemm_knobs.py 2>&1 | tail -1
