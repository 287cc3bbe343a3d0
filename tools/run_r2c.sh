#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gat_rgcn.py tests/test_gpu_trainer.py tests/test_gpu_nodeemb.py tests/test_gpu_original.py tests/test_gpu_pipeline.py -q > gpurun_out/r2c_rgcn.log 2>&1; tail -15 gpurun_out/r2c_rgcn.log
timeout 600 python -m pytest tests/test_gpu_fullsize_configs.py -q -k biokg > gpurun_out/r2c_biokg_full.log 2>&1; tail -5 gpurun_out/r2c_biokg_full.log
timeout 600 python tools/config_bench.py --epochs 20 > gpurun_out/r2c_configs.jsonl 2> gpurun_out/r2c_configs.err; cat gpurun_out/r2c_configs.jsonl; tail -3 gpurun_out/r2c_configs.err
GD_RGCN=transform timeout 600 python tools/config_bench.py --epochs 20 --configs biokg 2>/dev/null | tail -1
