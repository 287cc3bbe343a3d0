#!/bin/bash
{
python -m pytest tests/ -m gpu -x -q 2>&1 | tail -3
python tools/gemm_sweep.py
python bench.py --no-cpu-baseline
} > gpurun_out/gemm_check.log 2>&1
cat gpurun_out/gemm_check.log | cut -c1-1500
