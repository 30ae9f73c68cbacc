#!/bin/bash
# usage (under gpurun, one GPU): bash tools/r02_ncu_walk.sh <tag> — ncu --set full of the shipped walk kernel at N=1M, then 10M
TAG=${1:-r02g}
mkdir -p gpurun_out
KDNB_NO_GRAPH=1 timeout 28 ncu --set full --clock-control none --import-source on -k regex:walk2_kernel -s 3 -c 1 -f -o gpurun_out/walk_1M_${TAG} \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-10m > gpurun_out/ncu_walk_${TAG}.log 2>&1
echo "ncu walk 1M rc=$?"
KDNB_NO_GRAPH=1 timeout 42 ncu --set full --clock-control none --import-source on -k regex:walk2_kernel -s 3 -c 1 -f -o gpurun_out/walk_10M_${TAG} \
    python bench.py --steps 1 --warmup 3 --number 10000000 --no-cpu --no-10m > gpurun_out/ncu_walk10_${TAG}.log 2>&1
echo "ncu walk 10M rc=$?"
