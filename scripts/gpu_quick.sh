#!/usr/bin/env bash
# Quick perf visit: GPU tests, the three workloads, one ncu capture of the FIR kernel.
set -u
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -8 | tee $OUT/pytest_$TAG.log
for WL in ${WLS:-C3 C4 C5}; do
  timeout 300 python bench.py --workload $WL --steps 50 --warmup 5 --no-cpu-baseline --min-seconds 0.3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$WL', d['config']['kernel'], 'us/step %.2f fp32 frac %.3f hbm frac %.3f value %.0f e2e %.0f Msamp/s' % (d['ms_per_step']*1e3, d['roofline']['frac'], d['roofline']['hbm']['frac'], d['value'], d['e2e']['value']))
    elif 'Error' in l or 'error' in l: print(l.strip())
"
done
if [ "${NCU:-1}" = "1" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:stream_fir -s 10 -c 1 -f -o $OUT/prof_stream_$TAG \
  python bench.py --workload ${NCU_WL:-C3} --steps 20 --warmup 5 --no-cpu-baseline --min-seconds 0.001 > /dev/null 2>&1
ls -la $OUT/prof_stream_$TAG.ncu-rep
fi
