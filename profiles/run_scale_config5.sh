#!/bin/bash
# BASELINE config 5 (power-law 10M nodes / 200M edges, row-partitioned) at 1/2/4/8 GPUs of one box.
# usage: bash profiles/run_scale_config5.sh [scale]      (run under: gpurun --gpus 8)
SCALE=${1:-1.0}
mkdir -p gpurun_out
OUT=gpurun_out/scale_config5.jsonl
: > $OUT
timeout 900 python bench.py --workload powerlaw10m --scale $SCALE --steps 5 --warmup 3 2>gpurun_out/scale_n1.err | tail -1 >> $OUT
for N in 2 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) \
      bench.py --gpus $N --workload powerlaw10m --scale $SCALE --steps 5 --warmup 3 2>gpurun_out/scale_n$N.err | tail -1 >> $OUT
done
python - <<'PY'
import json
for l in open('gpurun_out/scale_config5.jsonl'):
    try:
        d=json.loads(l); print(d['n_gpus'], round(d['value'],3), 'epochs/s', round(d['ms_per_step'],2), 'ms', 'setup', round(d.get('setup_s',0),1), 's')
    except Exception as e:
        print('bad line', l[:200])
PY
tail -3 gpurun_out/scale_n8.err
