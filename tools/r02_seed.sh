#!/bin/bash
# usage (under gpurun, one GPU): bash tools/r02_seed.sh — the walk's traversal seeded with the 32 nodes of depth 5 (one
# vote over the 31 nodes above them) against the start at the root (KDNB_WALK_SEED=0); GPU tests with both, smoke()
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -1
KDNB_WALK_SEED=0 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "walk" 2>&1 | tail -1
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
bash tools/ab.sh walk_seed "1000000 10" "10000000 5" "125000 20" -- - "KDNB_WALK_SEED=0"
