#!/bin/bash
# usage (under gpurun --gpus G): bash tools/r02_final_multi.sh G — multi-GPU parity check at a size where every rank
# launches the 28-CTAs-per-SM peer kernel, then the bench line the driver runs (N=1M with its n10m sub-record)
G=${1:-2}
mkdir -p gpurun_out
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
run 29531 tests/multigpu_check.py 3000000 3 2>&1 | grep -E "multigpu_check|Error|assert" | tail -2
run 29532 bench.py --gpus $G --steps 10 --warmup 3 > gpurun_out/final_bench_n$G.log 2>&1
grep '^{' gpurun_out/final_bench_n$G.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']; t=d['n10m']; u=t['stage_ms_per_step']
print('G=$G 1M %.4e p-steps/s %.3f ms/step build %.3f walk %.3f exch %.3f kick %.3f parity %s | 10M %.4e %.3f ms build %.3f walk %.3f exch %.3f' % (d['value'], d['ms_per_step'], s['build'], s['walk'], s['exchange'], s['kick'], d['parity']['equals_single_gpu'] and d['parity']['ranks_bit_identical'], t['value'], t['ms_per_step'], u['build'], u['walk'], u['exchange']))" || tail -5 gpurun_out/final_bench_n$G.log
