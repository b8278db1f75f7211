#!/usr/bin/env bash
# one-visit experiment: the persistent kernel fed by tensor-map TMA (raw PCM boxes converted in place)
set -u
cd "$(dirname "$0")/.."
export PYTHONPATH=$PWD
chk() { local label=$1; shift
  env "$@" timeout 300 python tests/resident_check.py ${ARGS} 2>&1 | tail -2 | sed "s/^/$label: /"
}
for ARGS in "C3x 1300 2 44100 48000 7 882 3" "C5x 700 2 96000 44100 10 1920 2" "C4x 600 1 48000 16000 10 960 2" "q0 200 2 44100 48000 0 441 2"; do
  chk "packed default nt" SPXB_UMMA_RESIDENT=1
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
for WL in C5 C4 C3; do
  run "v1 $WL" $B --workload $WL
  run "v4 $WL packed default nt" SPXB_UMMA_RESIDENT=1 $B --workload $WL
  for NT in 64 80 96; do
    run "v4 $WL packed nt$NT" SPXB_UMMA_RESIDENT=1 SPXB_UMMA_NT=$NT $B --workload $WL
  done
  run "v4 $WL dense nt64" SPXB_UMMA_RESIDENT=1 SPXB_UMMA_DENSE=1 SPXB_UMMA_NT=64 $B --workload $WL
done
