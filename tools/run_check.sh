#!/bin/bash
{
python -m pytest tests/ -m gpu -x -q 2>&1 | tail -3
python tools/gemm_sweep.py 2>&1 | grep -E "tiles/CTA +(1|8|16) "
GD_TC_DEBUG=30 python tools/gemm_sweep.py 2>&1 | grep -E "tiles/CTA +(1|16) " | sed 's/tn128.*//'
python bench.py --no-cpu-baseline
} > gpurun_out/check.log 2>&1
cat gpurun_out/check.log | cut -c1-700
