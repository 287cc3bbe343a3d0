#!/bin/bash
{
python -m pytest tests/ -m gpu -x -q 2>&1 | tail -3
python tools/spmm_bench.py
python tools/gemm_sweep.py 2>&1 | grep -E "tiles/CTA +(8|16) "
python bench.py --no-cpu-baseline
} > gpurun_out/check.log 2>&1
cat gpurun_out/check.log | cut -c1-1700
