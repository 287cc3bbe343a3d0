#!/bin/bash
# the whole GPU suite + smoke + bench, as the driver runs them
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/full_gputests.log 2>&1; tail -14 gpurun_out/full_gputests.log
python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 600 gpurun_out/bench_full.json; tail -3 gpurun_out/bench_full.err
