#!/usr/bin/env bash
# one-visit experiment: dense tap tiles with the record-free MMA issue path
set -u
cd "$(dirname "$0")/.."
run() { local label=$1; shift
  env "$@" 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); g = d['roofline']['tensor']['geometry']; print('$label us/step %.2f  nt %d stages %d smem %d tiles %d' % (d['ms_per_step']*1e3, g['nt'], g['stages'], g['smem_bytes'], g['tiles']))
"
}
B="timeout 200 python bench.py --kernel tensor --steps 192 --warmup 10 --no-cpu-baseline --no-also --lean --min-seconds 0.3"
for WL in ${WLS:-C5 C3 C4}; do
  run "v1 $WL" $B --workload $WL
  for NT in 48 64 80; do
    for XS in 2 3 4; do
      run "v3 $WL dense nt$NT xstages<=$XS" SPXB_UMMA_RESIDENT=1 SPXB_UMMA_DENSE=1 SPXB_UMMA_NT=$NT SPXB_UMMA2_XSTAGES=$XS $B --workload $WL
    done
  done
done
