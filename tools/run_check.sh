#!/bin/bash
{
python -m pytest tests/ -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline | python -c "
import sys, json
for l in sys.stdin:
    l=l.strip()
    if l.startswith('{'):
        d=json.loads(l); print('value',d['value'],'hoisted',d['value_hoisted'],'e2e',d['e2e']['value'],'sync',d['e2e']['value_step_synchronous']); print(d['roofline']['frac'], d['roofline']['kernel_ms'], d['roofline']['other_kernels'])
"
python bench.py --workload powerlaw10m --scale 0.25 --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('powerlaw@0.25 1 GPU', d['value'], 'epochs/s  setup_s', d['setup_s'])
"
} > gpurun_out/check.log 2>&1
cat gpurun_out/check.log
