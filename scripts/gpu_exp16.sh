#!/usr/bin/env bash
# one-visit experiment: the persistent kernel with the A operand (byte planes) in tensor memory
set -u
cd "$(dirname "$0")/.."
export PYTHONPATH=$PWD SPXB_UMMA_RESIDENT=1
for ARGS in "C5x 700 2 96000 44100 10 1920 3" "C4x 600 1 48000 16000 10 960 2" "C3x 1300 2 44100 48000 7 882 3" "q0 200 2 44100 48000 0 441 2"; do
  timeout 120 python tests/resident_check.py ${ARGS} 2>&1 | tail -1 | cut -c1-300
  SPXB_UMMA_NT=112 timeout 120 python tests/resident_check.py ${ARGS} 2>&1 | tail -1 | cut -c1-300
done
run() { local label=$1; shift
  env "$@" 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); g = d['roofline']['tensor']['geometry']; print('$label us/step %.2f  nt %d stages %d smem %d tiles %d' % (d['ms_per_step']*1e3, g['nt'], g['stages'], g['smem_bytes'], g['tiles']))
"
}
B="timeout 100 python bench.py --kernel tensor --steps 192 --warmup 10 --no-cpu-baseline --no-also --lean --min-seconds 0.3"
for WL in C5 C4 C3; do
  run "planes in shared memory $WL" SPXB_UMMA_ATMEM=0 $B --workload $WL
  run "planes in TMEM $WL" $B --workload $WL
done
run "planes in TMEM C5 nt96" SPXB_UMMA_NT=96 $B --workload C5
run "one tile per CTA C3" SPXB_UMMA_RESIDENT=0 $B --workload C3
run "planes in TMEM C3 again" $B --workload C3
run "one tile per CTA C3 again" SPXB_UMMA_RESIDENT=0 $B --workload C3
