#!/bin/bash
# full GPU suite + smoke + default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/ -x -q -m gpu --durations=5 > gpurun_out/r2j_gputests.log 2>&1; tail -12 gpurun_out/r2j_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_r2j.json 2> gpurun_out/bench_r2j.err; tail -3 gpurun_out/bench_r2j.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r2j.json').read().strip().splitlines()[-1])
print('value', d['value'], 'hoisted', d.get('value_hoisted'), 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], d['roofline']['kernel_ms'], 'gemm share', d['roofline']['gemm_share_of_epoch'])
for k, v in d['roofline']['other_kernels'].items(): print(' ', k, round(v['ms'] * 1e3, 1), 'us', round(v['frac'], 3))
print('dense_ni', json.dumps(d.get('dense_ni'))[:1500])
print('bf16', json.dumps(d.get('bf16_gather'))[:300])
print('cpu', d.get('cpu_baseline'))
PY
