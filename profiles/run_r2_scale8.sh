#!/bin/bash
# Round 2: bench.py --gpus 8 as the driver launches it (headline replicas + the row-partitioned config-5 block with its
# parity check and the one-GPU point), then the partitioned config alone with the NCCL all-gather exchange (A/B).
mkdir -p gpurun_out
N=${1:-8}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_r2_gpus$N.json 2> gpurun_out/bench_r2_gpus$N.err
tail -c 2500 gpurun_out/bench_r2_gpus$N.json; tail -3 gpurun_out/bench_r2_gpus$N.err
NCCL_DEBUG=INFO NCCL_DEBUG_FILE=gpurun_out/nccl_r2_%p.log timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
    --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus $N --workload powerlaw10m --steps 10 --warmup 3 --partition-exchange nccl \
    > gpurun_out/bench_r2_gpus${N}_nooverlap.json 2> gpurun_out/bench_r2_gpus${N}_nooverlap.err
tail -c 1500 gpurun_out/bench_r2_gpus${N}_nooverlap.json; tail -3 gpurun_out/bench_r2_gpus${N}_nooverlap.err
grep -h -E "NVLS|Connected all|nChannels|Using network|via P2P" gpurun_out/nccl_r2_*.log | sort | uniq -c | sort -rn | head -8
ls gpurun_out/nccl_r2_*.log | head -1 | xargs -I{} cp {} gpurun_out/nccl_r2_rank_sample.log; rm -f gpurun_out/nccl_r2_[0-9]*.log
