#!/bin/bash
mkdir -p gpurun_out
timeout 100 python tools/dxdw_bench.py > gpurun_out/r2p_knobs.log 2>&1; cat gpurun_out/r2p_knobs.log
timeout 200 ncu --set full --import-source on --clock-control none -k regex:gemm_dxdw -s 2 -c 1 -f -o gpurun_out/r2p_dxdw python tools/dxdw_bench.py --once > gpurun_out/r2p_ncu.log 2>&1; tail -2 gpurun_out/r2p_ncu.log
