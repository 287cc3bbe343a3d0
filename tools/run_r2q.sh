#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r2q.json 2> gpurun_out/bench_r2q.err; tail -3 gpurun_out/bench_r2q.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r2q.json').read().strip().splitlines()[-1])
print('value', d['value'], 'hoisted', d.get('value_hoisted'), 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], d['roofline']['kernel_ms'], 'gemm share', d['roofline']['gemm_share_of_epoch'])
for k, v in d['roofline']['other_kernels'].items(): print(' ', k, round(v['ms'] * 1e3, 1), 'us', round(v['frac'], 3))
PY
