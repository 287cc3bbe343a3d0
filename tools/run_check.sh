#!/bin/bash
{
python -m pytest tests/test_gpu_trainer.py tests/test_gpu_engine.py -x -q 2>&1 | tail -3
timeout 500 python tools/config_bench.py --epochs 1000 --configs cora
} > gpurun_out/check.log 2>&1
cat gpurun_out/check.log | cut -c1-1200
