#!/usr/bin/env bash
# K1 (tile table in params) / K2 (converters help the history slide) / K3 (pipelined epilogue):
# correctness first, then A/B against the previous build in one visit, then the timeline.
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 300 python scripts/gpu_tensor_check.py 2>&1 | tail -4
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -4 | tee $OUT/pytest_k123.log
ab() {  # label lib env...
  local label=$1 libf=$2; shift 2
  for WL in C3 C4 C5; do
    env "$@" SPXB_LIB_PATH=$PWD/$libf timeout 300 python bench.py --workload $WL --kernel tensor --steps 50 --warmup 5 --no-cpu-baseline --min-seconds 0.3 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('$label $WL us/step %.2f' % (d['ms_per_step']*1e3))
"
  done
}
for ROUND in 1 2; do
  ab base ab/lib_base.so X=1
  ab k123 ab/lib_k123.so X=1
  ab k123-noinline ab/lib_k123.so SPXB_UMMA_INLINE_TILES=0
  ab k123-nohelp ab/lib_k123.so SPXB_UMMA_SLIDE_HELP=0
done 2>&1 | tee $OUT/ab_k123.log
SPXB_UMMA_TRACE=1 timeout 300 python scripts/gpu_trace.py C3 C4 C5 2>&1 | tee $OUT/trace_k123.log
