#!/usr/bin/env bash
# one-visit experiment: is the TMA-fed kernel waiting for HBM? (L2 prefetch ahead of the ring; L2-resident input)
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
B="timeout 100 python bench.py --kernel tensor --steps 192 --warmup 10 --no-cpu-baseline --no-also --lean --min-seconds 0.3"
P=$PWD/node_speex_resampler_b200
for WL in C5 C4; do
  for V in "" _pf4 _pf8; do
    run "v4$V $WL default nt" SPXB_LIB_PATH=$P/libspeexb200$V.so $B --workload $WL
    run "v4$V $WL default nt, ring 1 (L2-resident input)" SPXB_LIB_PATH=$P/libspeexb200$V.so $B --workload $WL --ring 1
    run "v4$V $WL dense nt64" SPXB_LIB_PATH=$P/libspeexb200$V.so SPXB_UMMA_DENSE=1 SPXB_UMMA_NT=64 $B --workload $WL
  done
done
