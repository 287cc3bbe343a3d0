#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gat_rgcn.py tests/test_gpu_trainer.py tests/test_gpu_fullsize_configs.py -q -k "gat or pubmed" > gpurun_out/r2f_gat.log 2>&1; tail -12 gpurun_out/r2f_gat.log
timeout 600 python tools/config_bench.py --epochs 200 --configs pubmed 2>/dev/null | tail -1
