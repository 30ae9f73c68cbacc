#!/bin/bash
# usage (under gpurun, one GPU, ~100 s): bash tools_quick_check.sh <tag>
# The shortest round-end sanity of a rebuilt libkdnb.so: a cross-section of the GPU parity suite (tree bit-exact,
# walk decisions, kick/drift, trajectory, error codes, empty and degenerate inputs), one short default bench line,
# then an A/B of one walk launch knob on shard-sized grids (tools/ab_walk_env.py; AB_VAR / AB_VALUES).
TAG=${1:-quick}
mkdir -p gpurun_out
t0=$SECONDS
timeout 40 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -p no:cacheprovider \
  -k "empty_bodies or degenerate or error_codes or (build_padded and 2049) or (build_dense and 5001) or (walk_ring and 5000) or (production_kernel and ring) or kick_drift or (simple_sim_trajectory and 1000-100) or reference_api_mirror or graph_replay_equals" \
  > gpurun_out/pytest_quick_${TAG}.log 2>&1
echo "pytest rc=$? $((SECONDS - t0)) s"; tail -n 3 gpurun_out/pytest_quick_${TAG}.log
timeout 45 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_quick_${TAG}.log 2>&1
echo "bench rc=$? $((SECONDS - t0)) s"; grep '^{' gpurun_out/bench_quick_${TAG}.log | cut -c1-600
[ -n "$SKIP_AB" ] && exit 0
AB_SIZES=${AB_SIZES:-"125000 250000 1000000"} timeout 35 python tools/ab_walk_env.py ${AB_VAR:-KDNB_WALK_PF} ${AB_VALUES:-0 1 3} --out gpurun_out/ab_walk_${TAG}.txt 2>&1 | tail -n 8
echo "ab rc=$? $((SECONDS - t0)) s"
