#!/bin/bash
# usage (under gpurun --gpus 8): bash tools/r02_scale.sh <tag>
# The scaling lines the driver takes at round end (bench.py at N = 8, 4, 2: top-level 1M line + n10m sub-record + parity),
# the split sort forced at 1M x 8 (default: from 2M particles), and the BASELINE configs[4] sweep (N = 100M, theta x MAX_PARTS).
TAG=${1:-r02}
mkdir -p gpurun_out
t0=$SECONDS
run() { timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
summ() { grep '^{' $1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']; n=d.get('n10m')
print('  N=%d: 1M %.4e p-steps/s %.3f ms/step (build %.3f walk %.3f exch %.3f kick %.3f) e2e %.3e parity %s' % (d['n_gpus'], d['value'], d['ms_per_step'], s['build'], s['walk'], s['exchange'], s['kick'], d['e2e']['value'], d.get('parity')))
if n:
  s=n['stage_ms_per_step']; print('       10M %.4e p-steps/s %.3f ms/step (build %.3f walk %.3f exch %.3f kick %.3f) e2e %.3e walk frac %s' % (n['value'], n['ms_per_step'], s['build'], s['walk'], s['exchange'], s['kick'], n['e2e']['value'], n.get('roofline',{}).get('frac')))
" || tail -5 $1; }
for g in 8 4 2; do
  run $g 2951$g bench.py --gpus $g --steps 10 --warmup 3 > gpurun_out/scale_${TAG}_n$g.log 2>&1
  echo "bench N=$g rc=$? $((SECONDS - t0)) s"; summ gpurun_out/scale_${TAG}_n$g.log
done
KDNB_SORT_SPLIT=1 run 8 29520 bench.py --gpus 8 --steps 10 --warmup 3 --no-10m > gpurun_out/scale_${TAG}_split1M.log 2>&1
echo "KDNB_SORT_SPLIT=1 at 1M x 8 rc=$? $((SECONDS - t0)) s"; summ gpurun_out/scale_${TAG}_split1M.log
run 8 29540 benchmarks/sweep_100m.py --steps 2 > gpurun_out/sweep_100m_${TAG}.log 2>&1
echo "sweep rc=$? $((SECONDS - t0)) s"; grep '^{' gpurun_out/sweep_100m_${TAG}.log | cut -c1-400
tail -3 gpurun_out/sweep_100m_${TAG}.log | cut -c1-300
