#!/usr/bin/env bash
# A/B of kernel variants (tile table in params, FAST instantiation, pipelined epilogue, converters
# helping the history slide) and of the host pipeline depth, in one visit.
set -u
OUT=gpurun_out; mkdir -p $OUT
for V in v1 v4 v5; do
  SPXB_LIB_PATH=$PWD/ab/lib_$V.so timeout 300 python scripts/gpu_tensor_check.py 2>&1 | tail -1 | sed "s/^/$V /"
done
ab() {  # label lib env...
  local label=$1 libf=$2; shift 2
  for WL in C3 C4 C5; do
    env "$@" SPXB_LIB_PATH=$PWD/$libf timeout 300 python bench.py --workload $WL --kernel tensor --steps 50 --warmup 5 --no-cpu-baseline --min-seconds 0.3 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('$label $WL us/step %.2f e2e %.0f (%.2f of pcie)' % (d['ms_per_step']*1e3, d['e2e']['value'], d['e2e']['pcie']['e2e_frac_of_ceiling']))
"
  done
}
for ROUND in 1 2; do
  ab base ab/lib_base.so X=1
  ab v1 ab/lib_v1.so X=1
  ab v1-nofast ab/lib_v1.so SPXB_UMMA_FAST=0
  ab v1-noinline ab/lib_v1.so SPXB_UMMA_INLINE_TILES=0
  ab v4-epi ab/lib_v4.so X=1
  ab v5-help ab/lib_v5.so X=1
  ab v1-slots6 ab/lib_v1.so SPXB_PIPELINE_SLOTS=6
  ab v1-slots8 ab/lib_v1.so SPXB_PIPELINE_SLOTS=8
done 2>&1 | tee $OUT/ab_trip7.log
