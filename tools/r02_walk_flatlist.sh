#!/bin/bash
# usage (under gpurun, one GPU): bash tools/r02_walk_flatlist.sh — planar inputs: interaction list of 120 entries (the z
# array's bytes hold list blocks; shipped) against 96 (libkdnb_f96.so, -DKDNB_W2_LIST_FLAT=96)
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "walk" 2>&1 | tail -1
KDNB_LIB=$PWD/multilanguagekdtree_b200/libkdnb_f96.so bash tools/ab.sh walk_f96 "1000000 10" "10000000 5" "125000 20" -- - | tail -3
bash tools/ab.sh walk_f120 "1000000 10" "10000000 5" "125000 20" -- - | tail -3
