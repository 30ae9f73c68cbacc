#!/bin/bash
# usage (under gpurun, one GPU): bash tools/r02_sanitize.sh  -> gpurun_out/sanitizer_r02.txt
mkdir -p gpurun_out
OUT=gpurun_out/sanitizer_r02.txt
echo "# compute-sanitizer over tools/sanitize_workload.py (round 2, final kernels)" > $OUT
echo "## memcheck" >> $OUT
timeout 500 compute-sanitizer --tool memcheck python tools/sanitize_workload.py 2>&1 | grep -E "COMPUTE-SANITIZER|ERROR SUMMARY|Invalid|sanitizer workload|Error" | head -20 >> $OUT
echo "## racecheck" >> $OUT
timeout 700 compute-sanitizer --tool racecheck python tools/sanitize_workload.py 2>&1 | grep -E "COMPUTE-SANITIZER|RACECHECK SUMMARY|hazard|sanitizer workload|Error" | head -20 >> $OUT
cat $OUT
