#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_ni_tc_kernel -s 2 -c 1 -o gpurun_out/r2_dense_ni_tc -f python tools/dense_ni_bench.py > gpurun_out/r2i_ncu.log 2>&1; tail -3 gpurun_out/r2i_ncu.log
ls -la gpurun_out/r2_dense_ni_tc.ncu-rep
