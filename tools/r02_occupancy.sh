#!/bin/bash
# usage (under gpurun, one GPU): bash tools/r02_occupancy.sh — walk time against the CTAs per SM (unused dynamic shared
# memory limits the residency of the production kernel; library built with -DKDNB_WALK_AB as libkdnb_ab.so).
# CTAs per SM = 233472 / (7936 + 1024 + pad), capped at 24 by the registers: pad 1536 -> 22, 2560 -> 20, 3584 -> 18,
# 5632 -> 16, 9728 -> 12
export KDNB_LIB=$PWD/multilanguagekdtree_b200/libkdnb_ab.so
bash tools/ab.sh walk_occupancy_1M "1000000 10" -- - "KDNB_WALK_PAD=1536" "KDNB_WALK_PAD=2560" "KDNB_WALK_PAD=3584" "KDNB_WALK_PAD=5632" "KDNB_WALK_PAD=9728"
bash tools/ab.sh walk_occupancy_10M "10000000 5" -- - "KDNB_WALK_PAD=2560" "KDNB_WALK_PAD=5632"
