#!/bin/bash
{
python -m pytest tests/test_gpu_trainer.py -x -q 2>&1 | tail -5
timeout 500 python tools/config_bench.py --epochs 20
} > gpurun_out/check.log 2>&1
cat gpurun_out/check.log | cut -c1-1200
