#!/bin/bash
{ python -m pytest tests/test_gpu_gcn.py -x -q 2>&1 | tail -2
for d in 0 32 8 40; do echo "== GD_TC_DEBUG=$d"; GD_TC_DEBUG=$d python tools/gemm_sweep.py 2>&1 | grep -E "tiles/CTA +(1|16) " | sed 's/tn128.*//'; done; } > gpurun_out/gemm_dbg.log 2>&1
cat gpurun_out/gemm_dbg.log
