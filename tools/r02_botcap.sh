#!/bin/bash
# usage (under gpurun, one GPU): bash tools/r02_botcap.sh  — A/B of the bottom-kernel segment size (compile-time knob)
mkdir -p gpurun_out
bash tools/ab.sh botcap_default "1000000 10" "10000000 5" "100000 20" -- - | tail -3
KDNB_NVCC_EXTRA="-DKDNB_BOT_CAP=1024 -DKDNB_BOT_THREADS=256 -DKDNB_BOT_MINB=8" python -m multilanguagekdtree_b200.build --force > /dev/null 2>&1
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "build" 2>&1 | tail -1
bash tools/ab.sh botcap_1024_8 "1000000 10" "10000000 5" "100000 20" -- - | tail -3
KDNB_NVCC_EXTRA="-DKDNB_BOT_CAP=1024 -DKDNB_BOT_THREADS=256 -DKDNB_BOT_MINB=6" python -m multilanguagekdtree_b200.build --force > /dev/null 2>&1
bash tools/ab.sh botcap_1024_6 "1000000 10" "10000000 5" "100000 20" -- - | tail -3
