#!/usr/bin/env bash
set -u
for LIBF in "$@"; do
  for WL in C3 C4; do
    SPXB_LIB_PATH=$PWD/$LIBF timeout 300 python bench.py --workload $WL --steps 200 --warmup 5 --no-cpu-baseline --min-seconds 0.3 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('$LIBF $WL us/step %.2f e2e %.0f Msamp/s depth %d' % (d['ms_per_step']*1e3, d['e2e']['value'], d['e2e']['pipeline_depth']))
"
  done
done
