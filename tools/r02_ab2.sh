#!/bin/bash
# usage (under gpurun, one GPU): bash tools/r02_ab2.sh <tag>   (library built with -DKDNB_WALK_AB for the cfg=1 legs)
TAG=${1:-r02c}
mkdir -p gpurun_out
t0=$SECONDS
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_${TAG}.log 2>&1
echo "pytest rc=$? $((SECONDS - t0)) s"; tail -3 gpurun_out/pytest_gpu_${TAG}.log
for cfg in ${CFGS:-0 1}; do for n in ${SIZES:-1000000 100000 10000000}; do
  KDNB_WALK_CFG=$cfg timeout 200 python bench.py --steps 10 --warmup 3 --number $n --no-cpu > gpurun_out/bench_${TAG}_n${n}_cfg$cfg.log 2>&1
  echo "n=$n cfg=$cfg rc=$?"; grep '^{' gpurun_out/bench_${TAG}_n${n}_cfg$cfg.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']
print('  %.4e p-steps/s  %.4f ms/step  build %.4f walk %.4f kick %.4f  frac %.4f  launches %d e2e %.3e' % (d['value'], d['ms_per_step'], s['build'], s['walk'], s['kick'], d['roofline']['frac'], d['gpu_launches'], d['e2e']['value']))"
done; done
echo "bench done $((SECONDS - t0)) s"
KDNB_NO_GRAPH=1 timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu \
    > gpurun_out/bench_under_ncu_${TAG}.log 2>&1
echo "launches rc=$? $((SECONDS - t0)) s"
KDNB_NO_GRAPH=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:walk2_kernel -s 3 -c 1 -f -o gpurun_out/walk_${TAG} \
    python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_walk_${TAG}.log 2>&1
echo "ncu walk rc=$? $((SECONDS - t0)) s"
