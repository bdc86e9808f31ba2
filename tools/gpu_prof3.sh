#!/bin/bash
# usage: tools/gpu_prof3.sh <tag> — ncu --set full of the three non-dominant integrate kernels
TAG=${1:-x}
mkdir -p gpurun_out
for k in k_celltop_up k_celltop_down k_front; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 40 -c 2 -o gpurun_out/prof_${k}_$TAG python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
done
ls -la gpurun_out | tail -5
