#!/bin/bash
# usage (under gpurun, one GPU): bash tools/r02_botthreads.sh — threads per CTA of the bottom build kernel (1024-slot
# segments): 256 x 6 CTAs/SM (default: 888 resident CTAs, the 1024 segments of N = 1M take two waves), 192 x 8 (1184
# resident: one wave), 128 x 9.  Variant libraries built here with KDNB_BOT_THREADS / KDNB_BOT_MINB.
for v in "" _bt192 _bt128; do
  export KDNB_LIB=$PWD/multilanguagekdtree_b200/libkdnb$v.so
  [ -n "$v" ] && timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "build" 2>&1 | tail -1
  bash tools/ab.sh botthreads$v "1000000 10" "10000000 5" "100000 20" -- - | tail -3
done
