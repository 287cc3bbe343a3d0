#!/bin/bash
mkdir -p gpurun_out
timeout 100 python tools/dxdw_bench.py > gpurun_out/r2q_knobs.log 2>&1; cat gpurun_out/r2q_knobs.log
