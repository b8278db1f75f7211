#!/usr/bin/env bash
# tile-width sweep of the tensor kernel (SPXB_UMMA_NT forces nt)
set -u
for WL in C3 C4 C5; do
  for NT in 0 48 64 80 96 112 128; do
    SPXB_UMMA_NT=$NT timeout 300 python bench.py --workload $WL --kernel tensor --steps 50 --warmup 5 --no-cpu-baseline --no-also --min-seconds 0.2 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); g = d['roofline']['tensor']['geometry']
        print('$WL forced_nt=$NT us/step %.2f' % (d['ms_per_step']*1e3), g)
"
  done
done
