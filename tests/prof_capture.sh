#!/bin/bash
# Profiling passes behind profiles/ (run under gpurun; a number printed under ncu is never a bench value).
# usage: tests/prof_capture.sh TAG      (keeps everything it writes under 64 MiB so that gpurun copies it back)
set -u
TAG=${1:-r02}
OUT=gpurun_out
mkdir -p $OUT
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --bwt-reads 0"
S="python bench.py --reads 10000000 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --bwt-reads 0"
# 1. launch list of the C2 workload (pilot + warm-up step + digest step + profiled step + timed step)
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file $OUT/launches_$TAG.csv $B > $OUT/launches_$TAG.log 2>&1
# 2. DRAM traffic of every launch of the dominant kernel during the same command
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:radix_scatter -c 320 --csv --log-file $OUT/traffic_scatter_$TAG.csv $B > $OUT/traffic_$TAG.log 2>&1
# 3. full-set captures on the C2 shape at 10 M reads (small reports): the round-1 text pass, the sort kernels and the refinement of round 2
ncu --set full --clock-control none --import-source on -k 'regex:dedup_cached|lms_flags|tile_popc' -c 4 -o $OUT/prof_${TAG}_text $S > $OUT/prof_${TAG}_text.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:radix_hist|radix_scatter|ext_local_sort' -s 40 -c 4 -o $OUT/prof_${TAG}_sort $S > $OUT/prof_${TAG}_sort.log 2>&1
du -sh $OUT; ls -la $OUT | tail -12
