#!/usr/bin/env bash
# A/B of library builds in ONE visit (boxes differ run to run): usage gpu_ab.sh libA.so libB.so ...
set -u
for ROUND in 1 2; do
for LIBF in "$@"; do
  for WL in ${WLS:-C3 C4 C5}; do
    SPXB_LIB_PATH=$PWD/$LIBF timeout 300 python bench.py --workload $WL --kernel tensor --steps 200 --warmup 10 --no-cpu-baseline --no-also --lean --min-seconds 0.3 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('$LIBF round $ROUND $WL us/step %.2f' % (d['ms_per_step']*1e3))
"
  done
done
done
