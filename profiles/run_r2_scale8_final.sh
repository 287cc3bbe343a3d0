#!/bin/bash
# End of round 2: bench.py --gpus N as the driver launches it (headline replicas + the row-partitioned config-5 block with its
# parity check and the one-GPU point) with the final kernels.
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r2_final_gpus$N.json 2> gpurun_out/bench_r2_final_gpus$N.err
tail -2 gpurun_out/bench_r2_final_gpus$N.err
python - <<PY
import json
d = json.loads(open('gpurun_out/bench_r2_final_gpus$N.json').read().strip().splitlines()[-1])
p = d.get('partitioned', {})
print('value', d['value'], 'n_gpus', d['n_gpus'])
print({k: p.get(k) for k in ('value', 'ms_per_step', 'efficiency', 'one_gpu_value', 'comm_nranks_seen', 'error')})
print(p.get('parity'))
print({k: v for k, v in p.items() if 'ms' in k or 'exchange' in k or 'halo' in k})
PY
