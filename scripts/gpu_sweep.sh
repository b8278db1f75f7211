#!/usr/bin/env bash
# Shape sweep of the streaming kernel on the three workloads.
set -u
TAG=${1:-s}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -4 | tee $OUT/pytest_$TAG.log
for WL in C3 C4 C5; do
for SH in ${SHAPES:-0 84 82 44 42}; do
  SPXB_STREAM_SHAPE=$SH timeout 300 python bench.py --workload $WL --steps 30 --warmup 5 --no-cpu-baseline --min-seconds 0.2 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$WL shape=$SH', d['config']['kernel'], 'us/step %.2f fp32 frac %.3f value %.0f e2e %.0f Msamp/s' % (d['ms_per_step']*1e3, d['roofline']['frac'], d['value'], d['e2e']['value']))
    elif 'Error' in l or 'error' in l: print(l.strip())
"
done
done
