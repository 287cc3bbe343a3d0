#!/bin/bash
# builds the two microbenchmarks if needed (the binaries are not tracked), then sweeps them
mkdir -p gpurun_out
for t in gather_bench gather_bench2; do
  [ -x tools/$t ] || nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/$t tools/$t.cu -lcuda || exit 1
done
B=tools/gather_bench2
{
timeout 60 tools/gather_bench | head -12
for pl in 0 1; do
  timeout 60 $B 0 64 235368 $pl
  timeout 60 $B 0 128 235368 $pl
done
for f in 64 128; do
  for cfg in "2 4 32" "3 4 32" "2 8 32" "3 8 32" "2 16 16" "3 16 16" "4 12 16" "2 12 32"; do
     set -- $cfg
     timeout 60 $B 2 $f 235368 1 $1 $2 1 $3
  done
  timeout 60 $B 1 $f 235368 1 2 8 1 32
  timeout 60 $B 1 $f 235368 1 3 16 1 16
done
timeout 60 $B 2 64 8000000 0 3 8 1 32
timeout 60 $B 0 64 8000000 0
} > gpurun_out/gather2.log 2>&1
tail -3 gpurun_out/gather2.log
