#!/bin/bash
# Profiling passes behind profiles/ (run under gpurun; a number printed under ncu is never a bench value).
# usage: tests/prof_capture.sh TAG      (keeps everything it writes under 64 MiB so that gpurun copies it back)
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
S="python bench.py --reads 10000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
# 1. launch list of the C2 workload (pilot + warm-up step + timed step)
ncu --metrics gpu__time_duration.sum --clock-control none -c 7000 --csv --log-file $OUT/launches_$TAG.csv $B > $OUT/launches_$TAG.log 2>&1
# 2. DRAM traffic of every launch of the dominant kernel during the same command
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:radix_scatter -c 216 --csv --log-file $OUT/traffic_scatter_$TAG.csv $B > $OUT/traffic_$TAG.log 2>&1
# 3. full-set captures on the C2 shape at 10 M reads (small reports): group-stage / table kernels of rounds 1-2, then the sort kernels of round 2
ncu --set full --clock-control none --import-source on -k 'regex:group_reduce|phrase_insert|rules_kernel|group_apply|dict_gather|full_apply' -c 11 -o $OUT/prof_${TAG}_dict $S > $OUT/prof_${TAG}_dict.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:radix_hist|radix_scatter' -s 70 -c 3 -o $OUT/prof_${TAG}_sort $S > $OUT/prof_${TAG}_sort.log 2>&1
du -sh $OUT; ls -la $OUT | tail -12
