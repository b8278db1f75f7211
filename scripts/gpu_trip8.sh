#!/usr/bin/env bash
# Where does the mono FAST instantiation lose its time on C4? Timelines of three builds.
set -u
OUT=gpurun_out; mkdir -p $OUT
for V in "v1 1" "v1 0" "v5 1"; do
  set -- $V
  echo "#### lib_$1 FAST=$2"
  SPXB_UMMA_FAST=$2 SPXB_LIB_PATH=$PWD/ab/lib_$1.so SPXB_UMMA_TRACE=1 timeout 300 python scripts/gpu_trace.py C4 2>&1
done | tee $OUT/trace_trip8.log
