#!/bin/bash
# Profiling passes behind profiles/ (run under gpurun; a number printed under ncu is never a bench value).
# usage: tests/prof_capture.sh TAG
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
# 1. launch list of ~2 steps of the C2 workload
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $OUT/launches_$TAG.csv $B > $OUT/launches_$TAG.log 2>&1
# 2. DRAM traffic of every launch of the dominant kernel during the same command
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:radix_scatter -c 210 --csv --log-file $OUT/traffic_scatter_$TAG.csv $B > $OUT/traffic_$TAG.log 2>&1
# 3. full-set capture of the top kernels on the C2 shape at 10 M reads (keeps the report small)
ncu --set full --clock-control none --import-source on -k 'regex:radix_scatter|group_apply|group_reduce|phrase_insert|dict_gather|rules_kernel|lms_flags|dedup_cached' -s 40 -c 10 -o $OUT/prof_$TAG python bench.py --reads 10000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > $OUT/prof_$TAG.log 2>&1
ls -la $OUT | tail -8
