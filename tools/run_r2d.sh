#!/bin/bash
mkdir -p gpurun_out
GD_RGCN=transform timeout 600 python -m pytest tests/test_gpu_fullsize_configs.py -q -k biokg 2>&1 | tail -4
GD_RGCN=edge timeout 600 python -m pytest tests/test_gpu_fullsize_configs.py -q -k biokg 2>&1 | grep -E "Error|passed|failed" | tail -4
timeout 600 python -m pytest tests/test_gpu_saint.py -q 2>&1 | tail -12
