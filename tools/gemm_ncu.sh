#!/bin/bash
for d in 30 0; do
GD_TC_DEBUG=$d ncu --set full --clock-control none --import-source on -k regex:gemm_rows_tc_kernel -s 2 -c 1 -f -o gpurun_out/prof_gemm_dbg$d python tools/gemm_one.py > gpurun_out/prof_gemm_dbg$d.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
