#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_engine.py tests/test_gpu_fullsize.py tests/test_gpu_trainer.py tests/test_gpu_gat_rgcn.py -x -q -m gpu 2>&1 | tail -3
GD_LIB_TAG=wt timeout 120 python tools/epoch_ab.py 2>&1 | tail -1
GD_GEMM_ROWS=ring GD_LIB_TAG=ring timeout 120 python tools/epoch_ab.py 2>&1 | tail -1
GD_LIB_TAG=wt-again timeout 120 python tools/epoch_ab.py 2>&1 | tail -1
