#!/usr/bin/env bash
# ncu --set full capture of the FIR kernel: scripts/gpu_ncu.sh <tag> <workload> <shape>
set -u
TAG=${1:-n}; WL=${2:-C3}; SH=${3:-0}
OUT=gpurun_out; mkdir -p $OUT
SPXB_STREAM_SHAPE=$SH timeout 900 ncu --set full --clock-control none --import-source on -k regex:stream_fir -s 8 -c 1 -f -o $OUT/prof_${WL}_${SH}_$TAG \
  python bench.py --workload $WL --steps 10 --warmup 3 --no-cpu-baseline --min-seconds 0.001 > /dev/null 2>&1
ls -la $OUT/prof_${WL}_${SH}_$TAG.ncu-rep
