#!/bin/bash
# what the GPU box looks like (cores, RAM, scratch space, GPUs, NVLink): printed once per round into gpurun_out/
{
  echo "== cpu"; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Socket|NUMA node\(s\)"
  echo "== mem"; free -g | head -2
  echo "== disk"; df -h /tmp /dev/shm | cat
  echo "== gpu"; nvidia-smi --query-gpu=index,name,memory.total,clocks.max.sm,power.limit --format=csv
  nvidia-smi topo -m 2>/dev/null | head -12
  echo "== nccl"; ls /usr/lib/x86_64-linux-gnu/libnccl* 2>/dev/null
} 2>&1
