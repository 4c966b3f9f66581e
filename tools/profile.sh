#!/bin/bash
# Profiling recipe (B200_PROFILING.md), 1 GPU.  Outputs go to gpurun_out/ (scratch);
# summaries are copied into profiles/ by hand.
set -x
mkdir -p gpurun_out
TAG=${1:-r01}
# 1. every launch with its device time (cold-cache, serialised: compare SHARES)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --seq 65536 --steps 2 --warmup 3 --no-e2e --no-cpu \
    > gpurun_out/launches_${TAG}.out 2>&1
# 2. full capture of the two tile kernels (skip the warm-up launches)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:chunk_kernel -s 2 -c 2 \
    -o gpurun_out/prof_${TAG} -f python bench.py --seq 32768 --steps 1 --warmup 3 --no-e2e --no-cpu \
    > gpurun_out/prof_${TAG}.out 2>&1
ls -la gpurun_out
