#!/bin/bash
# Round-2 starting point (one GPU): GPU tests, launch list, ncu --set full of the aggregation (after the two-pass
# layer-1 split) and loss kernels, bench line.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2_base_gputests.log 2>&1; tail -15 gpurun_out/r2_base_gputests.log
bash profiles/prof_launches.sh
cp gpurun_out/launches.csv gpurun_out/r2_base_launches.csv
bash profiles/prof_full.sh spmm_batched_kernel r2_base_spmm_batched 4 5
bash profiles/prof_full.sh 'edge_loss_fwd_kernel|node_loss_kernel' r2_base_edge_loss 2 2
python bench.py > gpurun_out/bench_r2_base.json 2> gpurun_out/bench_r2_base.err
tail -c 900 gpurun_out/bench_r2_base.json
