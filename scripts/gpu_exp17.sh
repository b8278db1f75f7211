#!/usr/bin/env bash
# one-visit tuning of the planes-in-TMEM kernel: descriptor prefetch, tile widths
set -u
cd "$(dirname "$0")/.."
export PYTHONPATH=$PWD SPXB_UMMA_RESIDENT=1
run() { local label=$1; shift
  env "$@" 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); g = d['roofline']['tensor']['geometry']; print('$label us/step %.2f  nt %d stages %d smem %d tiles %d' % (d['ms_per_step']*1e3, g['nt'], g['stages'], g['smem_bytes'], g['tiles']))
"
}
B="timeout 40 python bench.py --kernel tensor --steps 192 --warmup 10 --no-cpu-baseline --no-also --lean --min-seconds 0.3"
P=$PWD/node_speex_resampler_b200
for WL in C3 C4 C5; do
  run "one tile per CTA $WL" SPXB_UMMA_RESIDENT=0 $B --workload $WL
  run "planes in TMEM $WL" $B --workload $WL
  run "planes in TMEM, descriptors prefetched $WL" SPXB_LIB_PATH=$P/libspeexb200_pfm.so $B --workload $WL
done
run "planes in TMEM C5 nt96" SPXB_UMMA_NT=96 $B --workload C5
run "planes in TMEM C5 nt80" SPXB_UMMA_NT=80 $B --workload C5
run "planes in TMEM C4 nt96" SPXB_UMMA_NT=96 $B --workload C4
run "planes in TMEM C4 nt112" SPXB_UMMA_NT=112 $B --workload C4
run "planes in TMEM C3 nt96" SPXB_UMMA_NT=96 $B --workload C3
run "planes in TMEM C3 nt80" SPXB_UMMA_NT=80 $B --workload C3
