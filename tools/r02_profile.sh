#!/bin/bash
# usage (under gpurun, one GPU): bash tools/r02_profile.sh <tag>
# Fresh evidence of the shipped kernels: bench lines (1M, 10M), launch list, ncu --set full of the walk kernel,
# heaviest-first order A/B at 10M.  Numbers printed by bench.py under ncu are never bench values.
TAG=${1:-r02a}
mkdir -p gpurun_out
t0=$SECONDS
timeout 120 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_${TAG}.log 2>&1
echo "bench rc=$? $((SECONDS - t0)) s"; grep '^{' gpurun_out/bench_${TAG}.log | cut -c1-900
timeout 200 python bench.py --steps 5 --warmup 3 --number 10000000 --no-cpu > gpurun_out/bench_10M_${TAG}.log 2>&1
echo "bench10M rc=$? $((SECONDS - t0)) s"; grep '^{' gpurun_out/bench_10M_${TAG}.log | cut -c1-900
AB_SIZES="10000000" timeout 200 python tools/ab_walk_env.py KDNB_WALK_LPT 0 1 --out gpurun_out/ab_walk_lpt10M_${TAG}.txt 2>&1 | tail -3
echo "ab rc=$? $((SECONDS - t0)) s"
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu \
    > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
echo "launches rc=$? $((SECONDS - t0)) s"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:walk2_kernel -s 3 -c 1 -f -o gpurun_out/walk_${TAG} \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_walk_${TAG}.log 2>&1
echo "ncu walk rc=$? $((SECONDS - t0)) s"
ls -la gpurun_out | tail -8
