#!/bin/bash
# usage (under gpurun, one GPU): bash tools/r02_minb_grid.sh — production walk at 24 / 28 CTAs per SM against the grid
# size (N / 32 CTAs): where does the 72-register build start to win?
bash tools/ab.sh minb_grid "125000 20" "250000 20" "500000 10" "1000000 10" "4000000 5" -- "KDNB_WALK_MINB=24" "KDNB_WALK_MINB=28" -
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1
