#!/bin/bash
# usage (under gpurun, one GPU): bash tools/r02_profile2.sh <tag>
# Evidence of the shipped kernels: GPU test suite, the default bench line, launch lists with DRAM bytes at N=1M and N=10M
# (plain launches: ncu cannot see kernel nodes of a graph that holds a conditional node), ncu --set full of the walk, the
# bottom build kernel and the level partition at N=10M.  Numbers printed by bench.py under ncu are never bench values.
TAG=${1:-r02}
mkdir -p gpurun_out
t0=$SECONDS
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_${TAG}.log 2>&1
echo "pytest rc=$? $((SECONDS - t0)) s"; tail -2 gpurun_out/pytest_gpu_${TAG}.log
timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_${TAG}.log 2>&1
echo "bench rc=$? $((SECONDS - t0)) s"; grep '^{' gpurun_out/bench_${TAG}.log | cut -c1-300
timeout 120 python bench.py --steps 20 --warmup 3 --number 100000 --no-cpu --no-10m > gpurun_out/bench_100k_${TAG}.log 2>&1
echo "bench100k rc=$? $((SECONDS - t0)) s"
for n in 1000000 10000000; do
  KDNB_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
      --log-file gpurun_out/launches_${n}_${TAG}.csv python bench.py --steps 2 --warmup 3 --number $n --no-cpu --no-10m \
      > gpurun_out/bench_under_ncu_${n}_${TAG}.log 2>&1
  echo "launches $n rc=$? $((SECONDS - t0)) s"
done
KDNB_NO_GRAPH=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:walk2_kernel -s 3 -c 1 -f -o gpurun_out/walk_1M_${TAG} \
    python bench.py --steps 1 --warmup 3 --no-cpu --no-10m > gpurun_out/ncu_walk_${TAG}.log 2>&1
echo "ncu walk 1M rc=$? $((SECONDS - t0)) s"
KDNB_NO_GRAPH=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:walk2_kernel -s 3 -c 1 -f -o gpurun_out/walk_10M_${TAG} \
    python bench.py --steps 1 --warmup 3 --number 10000000 --no-cpu --no-10m > gpurun_out/ncu_walk10_${TAG}.log 2>&1
echo "ncu walk 10M rc=$? $((SECONDS - t0)) s"
KDNB_NO_GRAPH=1 timeout 300 ncu --set full --clock-control none -k regex:'build_bottom|level_partition|sort_downsweep|kick_drift' -s 60 -c 20 -f \
    -o gpurun_out/build_10M_${TAG} python bench.py --steps 1 --warmup 3 --number 10000000 --no-cpu --no-10m > gpurun_out/ncu_build_${TAG}.log 2>&1
echo "ncu build 10M rc=$? $((SECONDS - t0)) s"
ls -la gpurun_out | grep ${TAG} | tail -12
