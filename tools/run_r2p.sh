#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tn_wt_kernel -s 4 -c 2 -o gpurun_out/r2_gemm_tn_wt -f python tools/gemm_ab.py > gpurun_out/r2p_ncu.log 2>&1; tail -2 gpurun_out/r2p_ncu.log
ls -la gpurun_out/r2_gemm_tn_wt.ncu-rep
