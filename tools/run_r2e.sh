#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gat_rgcn.py tests/test_gpu_trainer.py tests/test_gpu_nodeemb.py -q > gpurun_out/r2e_rgcn.log 2>&1; tail -8 gpurun_out/r2e_rgcn.log
timeout 900 python -m pytest tests/test_gpu_fullsize_configs.py -q -k biokg 2>&1 | tail -4
python tools/rgcn_debug.py 1.0 2>&1 | grep -E "edge|err" | tail -12
timeout 900 python bench.py --workload biokg --steps 20 --warmup 3 > gpurun_out/bench_r2e_biokg.json 2> gpurun_out/bench_r2e_biokg.err; tail -c 2500 gpurun_out/bench_r2e_biokg.json; tail -5 gpurun_out/bench_r2e_biokg.err
timeout 600 python tools/config_bench.py --epochs 50 --configs pubmed,biokg 2>/dev/null | tail -2
