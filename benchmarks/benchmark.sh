#!/bin/bash
# The B200 counterpart of Parallel/RustVersion/benchmark.sh:1-16 (and of Parallel/{C,Cpp}Version/benchmark.sh): same
# particle counts, same 7 repetitions of a 10-step run, same times.txt layout ("<particles> <threads>" header line, then
# bash `time` blocks), so the reference's own post-processing reads it.  The thread-count axis of the reference is the
# GPU-count axis here (1 GPU per line; multi-GPU runs go through bench.py under torchrun).
#   usage (on a GPU box): bash benchmarks/benchmark.sh [out=times.txt]
out=${1:-times.txt}
here=$(cd "$(dirname "$0")/.." && pwd)
sim=$here/multilanguagekdtree_b200/kdtree-sim
particle_counts=(100000 1000000)
gpu_counts=(1)

rm -f "$out"
for parts in "${particle_counts[@]}"
do
	for gpus in "${gpu_counts[@]}"
	do
		echo $parts $gpus >> "$out"
		for cnt in {1..7}
		do
			{ time "$sim" --steps 10 --number $parts > /dev/null ; } 2>> "$out"
		done
	done
done
