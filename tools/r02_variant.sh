#!/bin/bash
# usage (under gpurun, one GPU): bash tools/r02_variant.sh <tag> <variant library suffix> [pytest -k expression]
# A/B of a compile-time variant built here as multilanguagekdtree_b200/libkdnb<suffix>.so against the default library:
# a parity subset with the variant, then N = 1M / 10M / 125k bench lines for both.
TAG=$1; SUF=$2; KEXPR=${3:-walk}
KDNB_LIB=$PWD/multilanguagekdtree_b200/libkdnb$SUF.so timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$KEXPR" 2>&1 | tail -1
bash tools/ab.sh ${TAG}_default "1000000 10" "10000000 5" "125000 20" -- - | tail -3
KDNB_LIB=$PWD/multilanguagekdtree_b200/libkdnb$SUF.so bash tools/ab.sh ${TAG}_variant "1000000 10" "10000000 5" "125000 20" -- - | tail -3
