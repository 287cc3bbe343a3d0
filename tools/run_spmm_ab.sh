#!/bin/bash
mkdir -p gpurun_out
{
python -m pytest tests/test_gpu_gcn.py tests/test_gpu_engine.py tests/test_gpu_golden.py -x -q 2>&1 | tail -5
GD_SPMM=pipe python tools/spmm_bench.py
python tools/spmm_bench.py
GD_SPMM_OVERSUB=2 python tools/spmm_bench.py
GD_SPMM_OVERSUB=4 python tools/spmm_bench.py
python bench.py --no-cpu-baseline
} > gpurun_out/spmm_ab.log 2>&1
tail -30 gpurun_out/spmm_ab.log
