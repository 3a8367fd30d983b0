#!/bin/bash
# full-set ncu captures of the text-pass and group-stage kernels (C2 shape at 10 M reads): rounds 1 and 2 of the first step
TAG=${1:-r01d}
OUT=gpurun_out
B="python bench.py --reads 10000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k 'regex:group_reduce|phrase_insert|rules_kernel|dedup_cached|lms_flags|rewrite_kernel' -c 12 -o $OUT/prof_${TAG}_misc $B > $OUT/prof_${TAG}_misc.log 2>&1
ls -la $OUT/*.ncu-rep
