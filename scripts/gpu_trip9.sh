#!/usr/bin/env bash
# ncu --set full with source correlation of the tensor kernel (C3, C5), current build
set -u
OUT=gpurun_out; mkdir -p $OUT
for WL in C5 C3; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:umma_fir -s 8 -c 1 -f -o $OUT/prof_umma_${WL}_v5 \
    python bench.py --workload $WL --kernel tensor --steps 10 --warmup 3 --no-cpu-baseline --min-seconds 0.001 > /dev/null 2>&1
done
ls -la $OUT/*.ncu-rep
