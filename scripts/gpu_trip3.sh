#!/usr/bin/env bash
# Tensor-kernel analysis visit: in-kernel timeline + ncu full capture.
set -u
TAG=${1:-t3}
OUT=gpurun_out
mkdir -p $OUT
SPXB_UMMA_TRACE=1 timeout 300 python scripts/gpu_trace.py C3 C4 C5 2>&1 | tee $OUT/trace_$TAG.log
if [ "${2:-ncu}" = "ncu" ]; then
for WL in C3 C5; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:umma_fir -s 8 -c 1 -f -o $OUT/prof_umma_${WL}_$TAG \
    python bench.py --workload $WL --kernel tensor --steps 10 --warmup 3 --no-cpu-baseline --min-seconds 0.001 > /dev/null 2>&1
done
fi
ls -la $OUT | tail -8
