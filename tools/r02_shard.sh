#!/bin/bash
# usage (under gpurun --gpus G): bash tools/r02_shard.sh G  — parity + A/B of the sharded tree build on G GPUs
G=${1:-2}
mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
for sb in 1 0; do
  KDNB_SHARD_BUILD=$sb run 29531 tests/multigpu_check.py 3000000 3 2>&1 | grep -E "multigpu_check|Error|assert" | tail -2
done
for n in 1000000 10000000; do for sb in 0 1; do
  KDNB_SHARD_BUILD=$sb run 29532 bench.py --gpus $G --steps 10 --warmup 3 --number $n --no-10m > gpurun_out/shard_g${G}_n${n}_sb$sb.log 2>&1
  grep '^{' gpurun_out/shard_g${G}_n${n}_sb$sb.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']
print('G=$G N=$n shard=$sb  %.4e p-steps/s  %.3f ms/step  build %.3f walk %.3f exch %.3f kick %.3f  parity %s' % (d['value'], d['ms_per_step'], s['build'], s['walk'], s['exchange'], s['kick'], d['parity']['equals_single_gpu'] and d['parity']['ranks_bit_identical']))" || tail -5 gpurun_out/shard_g${G}_n${n}_sb$sb.log
done; done
