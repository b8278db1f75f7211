#!/usr/bin/env bash
# one-visit experiment: does the L1 carve-out (shared memory left to L1) bind the converters' loads?
set -u
P=node_speex_resampler_b200
run() { local label=$1; shift
  env "$@" 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); g = d['roofline']['tensor']['geometry']; print('$label us/step %.2f  nt %d stages %d smem %d' % (d['ms_per_step']*1e3, g['nt'], g['stages'], g['smem_bytes']))
"
}
B="timeout 200 python bench.py --kernel tensor --steps 192 --warmup 10 --no-cpu-baseline --no-also --lean --min-seconds 0.3"
for WL in C5 C4; do
  run "v1 $WL" SPXB_UMMA_RESIDENT=0 $B --workload $WL
  run "v1 $WL smem padded to max" SPXB_UMMA_RESIDENT=0 SPXB_UMMA_PAD_SMEM=100000 $B --workload $WL
  run "v2 $WL" $B --workload $WL
  run "v2 $WL no-allocate loads" SPXB_LIB_PATH=$PWD/$P/lib_na.so $B --workload $WL
  run "v2 $WL xstages 2" SPXB_UMMA2_XSTAGES=2 $B --workload $WL
  run "v2 $WL nt 64 xstages 3" SPXB_UMMA_NT=64 SPXB_UMMA2_XSTAGES=3 $B --workload $WL
  run "v2 $WL nt 64 xstages 6" SPXB_UMMA_NT=64 $B --workload $WL
  run "v2 $WL nt 64 xstages 3 no-allocate" SPXB_LIB_PATH=$PWD/$P/lib_na.so SPXB_UMMA_NT=64 SPXB_UMMA2_XSTAGES=3 $B --workload $WL
done
