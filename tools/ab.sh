#!/bin/bash
# usage (under gpurun, one GPU): bash tools/ab.sh <tag> "<N> <steps>" ... -- "VAR=val VAR2=val" ...
# A/B timing of build-time knobs: for every size and every environment set, one bench line (--no-cpu) reduced to
# value, ms/step and the stage times.  "-" stands for the default environment.  Output: gpurun_out/ab_<tag>.txt
TAG=$1; shift
SIZES=()
while [ $# -gt 0 ] && [ "$1" != "--" ]; do SIZES+=("$1"); shift; done
shift
ENVSETS=("$@")
mkdir -p gpurun_out
OUT=gpurun_out/ab_${TAG}.txt
: > $OUT
for sz in "${SIZES[@]}"; do
  read N K <<< "$sz"
  for envs in "${ENVSETS[@]}"; do
    [ "$envs" = "-" ] && e="" || e="$envs"
    env $e timeout 200 python bench.py --steps $K --warmup 3 --number $N --no-cpu --no-10m > gpurun_out/ab_tmp.log 2>&1
    grep '^{' gpurun_out/ab_tmp.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); s=d['stage_ms_per_step']
print('N=$N [$envs] %.4e p-steps/s  %.3f ms/step  build %.3f walk %.3f kick %.3f  e2e %.3e' % (d['value'], d['ms_per_step'], s['build'], s['walk'], s['kick'], d['e2e']['value']))" >> $OUT 2>&1 || { echo "N=$N [$envs] FAILED" >> $OUT; tail -3 gpurun_out/ab_tmp.log >> $OUT; }
  done
done
cat $OUT
