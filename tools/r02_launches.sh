#!/bin/bash
# usage (under gpurun, one GPU): bash tools/r02_launches.sh <tag> — launch lists (device time + DRAM bytes) of one step at N = 1M, 10M, 100M
TAG=${1:-r02g}
mkdir -p gpurun_out
for n in 1000000 10000000; do
  KDNB_NO_GRAPH=1 timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
      --log-file gpurun_out/launches_${n}_${TAG}.csv python bench.py --steps 2 --warmup 3 --number $n --no-cpu --no-10m \
      > gpurun_out/bench_under_ncu_${n}_${TAG}.log 2>&1
  echo "launches $n rc=$?"
done
cat > /tmp/one100m.py <<'PY'
import sys; sys.path.insert(0, ".")
import multilanguagekdtree_b200 as kd
sim = kd.KDTreeSim()
sim.upload(kd.circular_orbits(100_000_000, seed=12345))
sim.simple_sim(1e-3, 2)
sim.synchronize()
print("done")
PY
KDNB_NO_GRAPH=1 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/launches_100000000_${TAG}.csv python /tmp/one100m.py > gpurun_out/ncu_100m_${TAG}.log 2>&1
echo "launches 100M rc=$?"; tail -2 gpurun_out/ncu_100m_${TAG}.log
