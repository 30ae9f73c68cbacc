#!/bin/bash
# usage (under gpurun --gpus 8): bash tools_scale.sh  -> gpurun_out/scale_*.log
mkdir -p gpurun_out
for g in 4 8; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 2951$g bench.py --gpus $g --steps 10 --no-cpu > gpurun_out/scale_1M_$g.log 2>&1
  grep '^{' gpurun_out/scale_1M_$g.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('1M GPUS', d['n_gpus'], '%.3e'%d['value'], d['stage_ms_per_step'], 'e2e %.3e'%d['e2e']['value'])"
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 5 --number 10000000 --no-cpu > gpurun_out/scale_10M_8.log 2>&1
grep '^{' gpurun_out/scale_10M_8.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('10M GPUS', d['n_gpus'], '%.3e'%d['value'], d['stage_ms_per_step'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 tests/multigpu_check.py 300000 3 > gpurun_out/multigpu_check_8.log 2>&1; tail -1 gpurun_out/multigpu_check_8.log
