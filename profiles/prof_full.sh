#!/bin/bash
# ncu --set full of the top kernels of one epoch (one GPU, short command).
# usage: bash profiles/prof_full.sh <kernel-regex> <out-name> [skip] [count]
mkdir -p gpurun_out
REGEX=${1:-spmm_vec_kernel}
OUT=${2:-prof}
SKIP=${3:-4}
COUNT=${4:-3}
ncu --set full --clock-control none --import-source on -k regex:$REGEX -s $SKIP -c $COUNT -f -o gpurun_out/$OUT \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${OUT}.log 2>&1
ls -la gpurun_out/$OUT.ncu-rep
