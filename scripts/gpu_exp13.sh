#!/usr/bin/env bash
# one-visit check: both tensor kernels after the convergent barrier initialisation
set -u
cd "$(dirname "$0")/.."
export PYTHONPATH=$PWD
run() { local label=$1; shift
  env "$@" 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); g = d['roofline']['tensor']['geometry']; print('$label us/step %.2f  nt %d stages %d smem %d tiles %d' % (d['ms_per_step']*1e3, g['nt'], g['stages'], g['smem_bytes'], g['tiles']))
"
}
B="timeout 100 python bench.py --kernel tensor --steps 192 --warmup 10 --no-cpu-baseline --no-also --lean --min-seconds 0.3"
for WL in C3 C4 C5; do
  run "one-tile-per-CTA $WL" SPXB_UMMA_RESIDENT=0 $B --workload $WL
  run "persistent $WL" SPXB_UMMA_RESIDENT=1 $B --workload $WL
  run "default $WL" $B --workload $WL
done
SPXB_LIB_PATH=$PWD/node_speex_resampler_b200/libspeexb200_trace.so SPXB_UMMA_RESIDENT=1 timeout 120 python scripts/gpu_trace2.py C3 2>&1 | head -12
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
