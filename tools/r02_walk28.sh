#!/bin/bash
# usage (under gpurun, one GPU): bash tools/r02_walk28.sh — production walk at 28 CTAs per SM (72 registers, stack of 240
# entries so that 28 x (7296 + 1024) bytes of shared memory fit) against 24 (80 registers); libkdnb_ab.so is built with
# -DKDNB_WALK_AB -DKDNB_W2_STACK=240
export KDNB_LIB=$PWD/multilanguagekdtree_b200/libkdnb_ab.so
KDNB_WALK_MINB=28 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "walk" 2>&1 | tail -1
bash tools/ab.sh walk28_1M "1000000 10" "125000 20" -- - "KDNB_WALK_MINB=28"
bash tools/ab.sh walk28_10M "10000000 5" -- - "KDNB_WALK_MINB=28"
