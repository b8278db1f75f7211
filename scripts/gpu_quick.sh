#!/usr/bin/env bash
# Quick perf visit: GPU tests, C3 tile-shape sweep, one ncu capture of the tiled kernel.
set -u
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -6 | tee $OUT/pytest_$TAG.log
for TR in 0 2 4 8; do
  SPXB_TILED_TR=$TR timeout 300 python bench.py --workload C3 --steps 50 --warmup 5 --no-cpu-baseline --min-seconds 0.3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('C3 TR=$TR us/step %.2f fp32 frac %.3f e2e %.0f Msamp/s' % (d['ms_per_step']*1e3, d['roofline']['frac'], d['e2e']['value']))
    elif 'Error' in l or 'error' in l: print(l.strip())
"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tiled_fir -s 10 -c 1 -f -o $OUT/prof_tiled_$TAG \
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --min-seconds 0.001 > /dev/null 2>&1
ls -la $OUT/prof_tiled_$TAG.ncu-rep
