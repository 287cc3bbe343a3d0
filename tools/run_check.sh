#!/bin/bash
{
python -m pytest tests/ -m gpu -x -q 2>&1 | tail -5
python bench.py --no-cpu-baseline
} > gpurun_out/check.log 2>&1
cat gpurun_out/check.log | cut -c1-3200
