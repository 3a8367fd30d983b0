#!/bin/bash
# full-set ncu captures of the non-sort kernels (C2 shape at 10 M reads) and of radix_hist
TAG=${1:-r01d}
OUT=gpurun_out
B="python bench.py --reads 10000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
ncu --set full --clock-control none --import-source on -k 'regex:group_reduce|phrase_insert|rules_kernel|group_apply|dedup_cached|lms_flags|ext_keys|full_apply|dict_gather|rewrite_kernel' -c 12 -o $OUT/prof_${TAG}_misc $B > $OUT/prof_${TAG}_misc.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:radix_hist|radix_scatter' -s 60 -c 3 -o $OUT/prof_${TAG}_sort $B > $OUT/prof_${TAG}_sort.log 2>&1
ls -la $OUT/*.ncu-rep
