#!/usr/bin/env bash
# one-visit experiment: persistent kernel with dense tap tiles and nt = 80 (one MMA per plane and K step)
set -u
run() { local label=$1; shift
  env "$@" 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); g = d['roofline']['tensor']['geometry']; print('$label us/step %.2f  nt %d stages %d smem %d tiles %d' % (d['ms_per_step']*1e3, g['nt'], g['stages'], g['smem_bytes'], g['tiles']))
"
}
B="timeout 200 python bench.py --kernel tensor --steps 192 --warmup 10 --no-cpu-baseline --no-also --lean --min-seconds 0.3"
for WL in C5 C4; do
  run "v1 $WL" SPXB_UMMA_RESIDENT=0 $B --workload $WL
  run "v2 $WL packed" $B --workload $WL
  run "v2 $WL dense nt80" SPXB_UMMA_DENSE=1 SPXB_UMMA_NT=80 $B --workload $WL
  run "v2 $WL packed nt80" SPXB_UMMA_NT=80 $B --workload $WL
  run "v2 $WL dense nt64" SPXB_UMMA_DENSE=1 SPXB_UMMA_NT=64 $B --workload $WL
  run "v1 $WL nt80" SPXB_UMMA_RESIDENT=0 SPXB_UMMA_NT=80 $B --workload $WL
  run "v1 $WL nt64" SPXB_UMMA_RESIDENT=0 SPXB_UMMA_NT=64 $B --workload $WL
done
