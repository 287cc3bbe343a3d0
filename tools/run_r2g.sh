#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_gcn.py tests/test_gpu_engine.py tests/test_gpu_fullsize.py tests/test_gpu_trainer.py -x -q -m gpu > gpurun_out/r2g_tests.log 2>&1; tail -5 gpurun_out/r2g_tests.log
timeout 600 python bench.py > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r2g.json').read().strip().splitlines()[-1])
print('value', d['value'], 'hoisted', d.get('value_hoisted'), 'e2e', d['e2e']['value'], 'roof', d['roofline']['frac'], d['roofline']['kernel_ms'])
for k, v in d['roofline']['other_kernels'].items(): print(' ', k, round(v['ms'] * 1e3, 1), 'us', round(v['frac'], 3))
PY
timeout 600 python tools/gemm_knobs.py 2>&1 | tee gpurun_out/r2g_knobs.log | tail -12
timeout 300 python tools/dense_ni_bench.py 2>&1 | tee gpurun_out/r2g_dense_ni.log | tail -2
