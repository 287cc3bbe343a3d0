#!/bin/bash
# 2 GPUs: the partitioned engine's NCCL / symmetric-memory tests + bench.py --gpus 2 on a 1/10-scale config-5 graph
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_dist.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 \
    bench.py --gpus 2 --steps 20 --warmup 5 --partition-scale 0.1 > gpurun_out/bench_r2m_gpus2.json 2> gpurun_out/bench_r2m_gpus2.err
tail -3 gpurun_out/bench_r2m_gpus2.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_r2m_gpus2.json').read().strip().splitlines()[-1])
p = d.get('partitioned', {})
print('value', d['value'], 'n_gpus', d['n_gpus'], 'partitioned:', {k: p.get(k) for k in ('value', 'ms_per_step', 'efficiency', 'one_gpu_value', 'parity', 'comm_nranks_seen', 'error')})
PY
