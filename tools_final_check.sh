#!/bin/bash
# usage (under gpurun, one GPU): bash tools_final_check.sh <tag>
# Round-end evidence in priority order (every leg has its own timeout; outputs under gpurun_out/):
#  1) the GPU parity suite, 2) the default bench line (N=1M) and the N=100k line (BASELINE configs[1]),
#  3) the launch list of one N=10M step with DRAM bytes (where the replicated build's time goes at the north-star size),
#  4) ncu --set full of the heaviest build kernels at N=10M (one whole step's worth of their launches: 4 + 12 + 1).
# Numbers printed by bench.py under ncu are never bench values.
TAG=${1:-r01_final}
mkdir -p gpurun_out
t0=$SECONDS
timeout 420 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_${TAG}.log 2>&1
echo "pytest rc=$? $((SECONDS - t0)) s"; tail -3 gpurun_out/pytest_gpu_${TAG}.log
timeout 240 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}.log 2>&1
echo "bench rc=$? $((SECONDS - t0)) s"; grep '^{' gpurun_out/bench_${TAG}.log | cut -c1-400
timeout 120 python bench.py --steps 20 --warmup 3 --number 100000 --no-cpu > gpurun_out/bench_100k_${TAG}.log 2>&1
echo "bench100k rc=$? $((SECONDS - t0)) s"; grep '^{' gpurun_out/bench_100k_${TAG}.log | cut -c1-300
timeout 240 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_10M_${TAG}.csv python bench.py --steps 1 --warmup 3 --number 10000000 --no-cpu \
    > gpurun_out/bench_under_ncu_10M_${TAG}.log 2>&1
echo "launches10M rc=$? $((SECONDS - t0)) s"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'build_bottom|level_partition|sort_downsweep' -s 17 -c 17 -f \
    -o gpurun_out/build_10M_${TAG} python bench.py --steps 1 --warmup 3 --number 10000000 --no-cpu \
    > gpurun_out/ncu_build_10M_${TAG}.log 2>&1
echo "ncu build rc=$? $((SECONDS - t0)) s"
ls -la gpurun_out | tail -12
