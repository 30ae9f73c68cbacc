#!/bin/bash
# usage (under gpurun, one GPU): bash tools/r02_final.sh — what the driver runs at round end: the GPU tests, smoke(),
# both bench arms (the reference arm shortened)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/final_bench_n1.log 2>&1; grep '^{' gpurun_out/final_bench_n1.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('N=1M %.4e (%.3f ms) e2e %.3e frac %.3f launches %d cpu %.3e | 10M %.4e frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'], d['gpu_launches'], d['cpu_baseline']['value'], d['n10m']['value'], d['n10m']['roofline']['frac']))"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | grep '^{' | cut -c1-200
