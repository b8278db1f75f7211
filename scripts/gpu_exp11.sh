#!/usr/bin/env bash
# one-visit experiment: TMA-fed persistent kernel, two accumulator sets (nt <= 64) vs one
set -u
cd "$(dirname "$0")/.."
export PYTHONPATH=$PWD SPXB_UMMA_RESIDENT=1
chk() { local label=$1; shift
  env "$@" timeout 300 python tests/resident_check.py ${ARGS} 2>&1 | tail -1 | sed "s/^/$label: /"
}
for ARGS in "C3x 1300 2 44100 48000 7 882 3" "C5x 700 2 96000 44100 10 1920 2" "C4x 600 1 48000 16000 10 960 2" "q0 200 2 44100 48000 0 441 2"; do
  chk "dense nt64" SPXB_UMMA_DENSE=1 SPXB_UMMA_NT=64
  chk "packed nt48" SPXB_UMMA_NT=48
done
run() { local label=$1; shift
  env "$@" 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); g = d['roofline']['tensor']['geometry']; print('$label us/step %.2f  nt %d stages %d smem %d tiles %d' % (d['ms_per_step']*1e3, g['nt'], g['stages'], g['smem_bytes'], g['tiles']))
"
}
B="timeout 200 python bench.py --kernel tensor --steps 192 --warmup 10 --no-cpu-baseline --no-also --lean --min-seconds 0.3"
for WL in C5 C3 C4; do
  run "v1 $WL" SPXB_UMMA_RESIDENT=0 $B --workload $WL
  run "v4 $WL packed default nt" $B --workload $WL
  run "v4 $WL dense nt64" SPXB_UMMA_DENSE=1 SPXB_UMMA_NT=64 $B --workload $WL
  run "v4 $WL packed nt64" SPXB_UMMA_NT=64 $B --workload $WL
  run "v4 $WL dense nt48" SPXB_UMMA_DENSE=1 SPXB_UMMA_NT=48 $B --workload $WL
  run "v4 $WL dense nt64 xs4" SPXB_UMMA_DENSE=1 SPXB_UMMA_NT=64 SPXB_UMMA2_XSTAGES=4 $B --workload $WL
done
