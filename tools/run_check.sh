#!/bin/bash
{
python -m pytest tests/ -m gpu -x -q 2>&1 | tail -2
python bench.py --no-cpu-baseline | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('value',d['value'],'hoisted',d['value_hoisted'],'e2e',d['e2e']['value'],'sync',d['e2e']['value_step_synchronous']); print(d['roofline']['kernel_ms'], d['roofline']['other_kernels'])
"
} > gpurun_out/check.log 2>&1
cat gpurun_out/check.log
