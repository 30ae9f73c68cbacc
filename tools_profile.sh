#!/bin/bash
# usage (under gpurun): bash tools_profile.sh <tag> [N]
# 1) launch list of one bench run with per-launch device time and DRAM bytes (cold-cache, serialised: compare SHARES),
# 2) ncu --set full of the walk kernel.  Numbers printed by bench.py under ncu are never bench values.
TAG=${1:-r01}
N=${2:-1000000}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --number $N --no-cpu \
    > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:walk2?_kernel -s 3 -c 1 -f -o gpurun_out/walk_${TAG} \
    python bench.py --steps 1 --warmup 3 --number $N --no-cpu > gpurun_out/ncu_walk_${TAG}.log 2>&1
ls -la gpurun_out | tail -5
