#!/bin/bash
{
python -m pytest tests/test_gpu_gcn.py tests/test_gpu_engine.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -1
python tools/gemm_sweep.py 2>&1 | grep -E "tiles/CTA +(1|8|16) " | sed 's/tn128.*//'
GD_TC_DEBUG=30 python tools/gemm_sweep.py 2>&1 | grep -E "tiles/CTA +(16) " | sed 's/tn128.*//'
GD_TC_DEBUG=94 python tools/gemm_sweep.py 2>&1 | grep -E "tiles/CTA +(16) " | sed 's/tn128.*//'
} > gpurun_out/check.log 2>&1
cat gpurun_out/check.log
