#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_trainer.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python tools/config_bench.py --epochs 300 --configs cora,cora-dense 2>/dev/null | tee gpurun_out/r2l_configs.jsonl | cut -c1-700
