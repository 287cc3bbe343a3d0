#!/bin/bash
# Round-2 closing run with the chained dX1 -> dW_del1 kernel (gemm_dxdw_wt.cu) as the default: GPU parity suite, the bench
# line, one ncu --set full capture of the chained kernel, the launch list of the epoch.
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/r2_final3_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2_final3_tests.log
tail -3 gpurun_out/r2_final3_tests.log
timeout 300 python bench.py > gpurun_out/bench_r2_final3.json 2> gpurun_out/bench_r2_final3.err; tail -c 600 gpurun_out/bench_r2_final3.json
timeout 120 ncu --set full --clock-control none --import-source on -k regex:gemm_dxdw -s 2 -c 1 -f -o gpurun_out/r2_final3_dxdw \
    python tools/dxdw_bench.py --once > gpurun_out/r2_final3_dxdw.log 2>&1
ls -la gpurun_out/r2_final3_dxdw.ncu-rep
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r2_final3_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_final3_bench_under_ncu.log 2>&1
wc -l gpurun_out/r2_final3_launches.csv
