#!/bin/bash
# One profiling pass of the bench workload on one B200 (run under gpurun from the repo root):
#   1. the plain command must exit 0 first (numbers printed under ncu are never bench values);
#   2. launch list: per-launch gpu__time_duration of two steps (caches left warm);
#   3. `ncu --set full` of the three heaviest kernels of the 4th step, once with ncu's cache flush
#      (cold) and once with --cache-control none (warm).
# usage: bash profiles/capture.sh <tag>      -> gpurun_out/{launches,prof}_<tag>*.{csv,ncu-rep}
set -u
TAG=${1:-r01_v10}
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 400 --csv \
    --log-file gpurun_out/launches_${TAG}_warm.csv $CMD > gpurun_out/ncu_l_$TAG.log 2>&1
K='regex:scan_targets_kernel|confirm_pairs_kernel|build_keys_insert_kernel'
ncu --set full --clock-control none --import-source on -k "$K" --launch-skip 9 --launch-count 3 -f \
    -o gpurun_out/prof_${TAG}_cold $CMD > gpurun_out/ncu_c_$TAG.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k "$K" --launch-skip 9 --launch-count 3 -f \
    -o gpurun_out/prof_${TAG}_warm $CMD > gpurun_out/ncu_w_$TAG.log 2>&1
ls -la gpurun_out/*$TAG*
