#!/bin/bash
{
timeout 500 python tools/config_bench.py --epochs 1000 --configs pubmed,cora
} > gpurun_out/check.log 2>&1
cat gpurun_out/check.log | cut -c1-1500
