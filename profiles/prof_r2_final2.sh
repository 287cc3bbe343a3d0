#!/bin/bash
# Round-2 final profile set (one GPU, end of the round): launch list of the epoch, ncu --set full of the aggregation, the
# weights-in-TMEM GEMMs (gemm_rows_wt_kernel / gemm_tn_wt_kernel) and the tensor-core dense-block NI kernel.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_final2_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_final2_bench_under_ncu.log 2>&1
wc -l gpurun_out/r2_final2_launches.csv
full() {  # regex out skip count cmd...
  local re=$1 out=$2 skip=$3 cnt=$4; shift 4
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$re -s $skip -c $cnt -f -o gpurun_out/$out "$@" > gpurun_out/$out.log 2>&1
  ls -la gpurun_out/$out.ncu-rep
}
full spmm_batched_kernel r2_final2_spmm_batched 8 4 python bench.py --steps 2 --warmup 3 --no-cpu-baseline
full 'gemm_rows_wt_kernel|gemm_tn_wt_kernel' r2_final2_gemm 16 8 python bench.py --steps 2 --warmup 3 --no-cpu-baseline
full dense_ni_tc_kernel r2_final2_dense_ni_tc 2 1 python tools/dense_ni_bench.py
