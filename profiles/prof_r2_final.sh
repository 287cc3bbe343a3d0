#!/bin/bash
# Round-2 profile set (one GPU): launch list of the epoch, ncu --set full of the aggregation, the fused loss kernels, the
# tcgen05 GEMMs (Collab epoch) and the RGCN edge kernel (BioKG step).
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_final_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_under_ncu.log 2>&1
wc -l gpurun_out/r2_final_launches.csv
full() {  # regex out skip count cmd...
  local re=$1 out=$2 skip=$3 cnt=$4; shift 4
  ncu --set full --clock-control none --import-source on -k regex:$re -s $skip -c $cnt -f -o gpurun_out/$out "$@" > gpurun_out/$out.log 2>&1
  ls -la gpurun_out/$out.ncu-rep
}
full spmm_batched_kernel r2_final_spmm_batched 7 4 python bench.py --steps 2 --warmup 3 --no-cpu-baseline
full 'node_loss_kernel|edge_loss_fwd_kernel' r2_final_loss 2 2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline
full 'gemm_rows_tc_kernel|gemm_tn_tc_kernel' r2_final_gemm 11 8 python bench.py --steps 2 --warmup 3 --no-cpu-baseline
full 'rgcn_edge_kernel' r2_final_rgcn_edge 4 2 python bench.py --workload biokg --steps 2 --warmup 3 --no-cpu-baseline
