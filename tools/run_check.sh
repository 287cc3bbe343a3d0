#!/bin/bash
{
python -m pytest tests/ -m gpu -x -q 2>&1 | tail -3
python tools/gemm_sweep.py 2>&1 | grep -E "tiles/CTA +(1|8|16) "
python bench.py --no-cpu-baseline
for i in 1 2 3; do python -m pytest tests/test_gpu_gcn.py -x -q -k "gemm" 2>&1 | tail -1; done
} > gpurun_out/check.log 2>&1
cat gpurun_out/check.log | cut -c1-900
