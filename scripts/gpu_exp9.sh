#!/usr/bin/env bash
# one-visit experiment: the persistent kernel with its PCM loads and/or MMAs compiled out
# (what bounds a stage once the LDG path is out of the way?)
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
for V in "" _noload _nomma _skel; do
  L=$PWD/node_speex_resampler_b200/libspeexb200$V.so
  for WL in C5; do
    run "v3$V $WL packed default nt" SPXB_LIB_PATH=$L SPXB_UMMA_RESIDENT=1 $B --workload $WL
    run "v3$V $WL dense nt64 xs3" SPXB_LIB_PATH=$L SPXB_UMMA_RESIDENT=1 SPXB_UMMA_DENSE=1 SPXB_UMMA_NT=64 $B --workload $WL
  done
  true
done
