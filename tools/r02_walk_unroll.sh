#!/bin/bash
# usage (under gpurun, one GPU): bash tools/r02_walk_unroll.sh — drain loop of the walk with 2 (default) / 4 list blocks
# per iteration, and 2 blocks at 28 CTAs per SM (72 registers, -DKDNB_W2_STACK=240)
bash tools/ab.sh unroll_2 "1000000 10" "10000000 5" "125000 20" -- - | tail -3
KDNB_LIB=$PWD/multilanguagekdtree_b200/libkdnb_u4.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "walk" 2>&1 | tail -1
KDNB_LIB=$PWD/multilanguagekdtree_b200/libkdnb_u4.so bash tools/ab.sh unroll_4 "1000000 10" "10000000 5" "125000 20" -- - | tail -3
KDNB_LIB=$PWD/multilanguagekdtree_b200/libkdnb_s240.so bash tools/ab.sh unroll_2_minb28 "1000000 10" "10000000 5" "125000 20" -- "KDNB_WALK_MINB=28" | tail -3
