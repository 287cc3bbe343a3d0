#!/bin/bash
{
python -m pytest tests/ -m gpu -x -q 2>&1 | tail -3
python bench.py
python bench.py --impl reference --steps 3 --warmup 1
} > gpurun_out/check.log 2>&1
cat gpurun_out/check.log | cut -c1-3500
