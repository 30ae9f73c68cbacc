#!/bin/bash
# usage (under gpurun, one GPU): bash tools/r02_walk_l64.sh — interaction list of 64 entries (6144 bytes of shared
# memory: 32 CTAs per SM fit) at 28 / 32 CTAs per SM against the shipped 96 entries at 28
bash tools/ab.sh walk_l96 "1000000 10" "10000000 5" -- - | tail -2
KDNB_LIB=$PWD/multilanguagekdtree_b200/libkdnb_l64.so bash tools/ab.sh walk_l64 "1000000 10" "10000000 5" -- "KDNB_WALK_MINB=28" "KDNB_WALK_MINB=32" | tail -4
