#!/bin/bash
# usage (under gpurun --gpus G): bash tools/r02_split.sh G  — parity + A/B of the split sort on G GPUs
G=${1:-2}
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
KDNB_SORT_SPLIT=1 KDNB_SHARD_BUILD=1 run 29531 tests/multigpu_check.py 3000000 3 2>&1 | grep -E "multigpu_check|Error|assert" | tail -2
for n in ${SIZES:-1000000 10000000}; do for sp in 0 1; do
  KDNB_SORT_SPLIT=$sp run 29532 bench.py --gpus $G --steps 10 --warmup 3 --number $n --no-10m > gpurun_out/split_g${G}_n${n}_sp$sp.log 2>&1
  grep '^{' gpurun_out/split_g${G}_n${n}_sp$sp.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']
print('G=$G N=$n split=$sp  %.4e p-steps/s  %.3f ms/step  build %.3f walk %.3f exch %.3f kick %.3f  parity %s' % (d['value'], d['ms_per_step'], s['build'], s['walk'], s['exchange'], s['kick'], d['parity']['equals_single_gpu'] and d['parity']['ranks_bit_identical']))" || tail -5 gpurun_out/split_g${G}_n${n}_sp$sp.log
done; done
