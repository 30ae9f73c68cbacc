#!/bin/bash
# usage (under gpurun, one GPU): bash tools/r02_cpasync.sh — the walk with cp.async staging of the leaf particles
# (-DKDNB_WALK_CPASYNC) against the default (plain loads), parity first.
mkdir -p gpurun_out
OUT=gpurun_out/ab_walk_cpasync.txt
echo "# default build (leaf particles: LDG.256 into registers, STS into the list)" > $OUT
bash tools/ab.sh cp_default "1000000 20" "10000000 5" "100000 20" -- - | tail -3 >> $OUT
KDNB_NVCC_EXTRA="-DKDNB_WALK_CPASYNC" python -m multilanguagekdtree_b200.build --force > /dev/null 2>&1
echo "# -DKDNB_WALK_CPASYNC (leaf particles: cp.async global -> shared, 8 bytes per coordinate; LDGSTS in the SASS: $(cuobjdump -sass multilanguagekdtree_b200/libkdnb.so | grep -c LDGSTS))" >> $OUT
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "walk or simple_sim or golden" 2>&1 | tail -1 >> $OUT
bash tools/ab.sh cp_async "1000000 20" "10000000 5" "100000 20" -- - | tail -3 >> $OUT
cat $OUT
