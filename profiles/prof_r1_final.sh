#!/bin/bash
# Round-1 final profile set (one GPU): launch list, ncu --set full of the aggregation / GEMM / loss kernels, bench line.
mkdir -p gpurun_out
bash profiles/prof_launches.sh
cp gpurun_out/launches.csv gpurun_out/r1_final_launches.csv
bash profiles/prof_full.sh spmm_batched_kernel r1_final_spmm_batched 4 4
bash profiles/prof_full.sh "gemm_rows_tc_kernel|gemm_tn_tc_kernel" r1_final_gemm 8 8
bash profiles/prof_full.sh edge_loss_fwd_kernel r1_final_edge_loss 1 1
python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err
tail -c 600 gpurun_out/bench_r1_final.json
