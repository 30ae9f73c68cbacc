#!/bin/bash
# usage (under gpurun --gpus 8): bash tools_scale.sh  -> gpurun_out/scale_*.log
mkdir -p gpurun_out
run() { # name gpus extra-env number steps port
  env $3 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $2 --master-addr 127.0.0.1 --master-port $6 bench.py --gpus $2 --steps $5 --number $4 --no-cpu > gpurun_out/scale_$1.log 2>&1
  grep '^{' gpurun_out/scale_$1.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'GPUS', d['n_gpus'], '%.3e'%d['value'], 'ms/step %.3f'%d['ms_per_step'], d['stage_ms_per_step'], 'e2e %.3e'%d['e2e']['value'])"
}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 tests/multigpu_check.py 300000 3 > gpurun_out/multigpu_check_8.log 2>&1; tail -1 gpurun_out/multigpu_check_8.log
run 1M_8_p2p 8 KDNB_X=0 1000000 10 29518
run 1M_4_p2p 4 KDNB_X=0 1000000 10 29514
run 10M_8_p2p 8 KDNB_X=0 10000000 5 29520
