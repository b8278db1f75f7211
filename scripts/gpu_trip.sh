#!/usr/bin/env bash
# One GPU-box visit: tests, benches, ncu captures. Outputs land in gpurun_out/ (scratch).
# usage: scripts/gpu_trip.sh <tag> [tests|notests]
set -u
TAG=${1:-trip}
OUT=gpurun_out
mkdir -p $OUT
if [ "${2:-tests}" = "tests" ]; then
  timeout 1200 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -25 | tee $OUT/pytest_$TAG.log
fi
for WL in C3 C4 C5; do
  timeout 600 python bench.py --workload $WL --steps 50 --warmup 5 > $OUT/bench_${WL}_$TAG.json 2> $OUT/bench_${WL}_$TAG.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_${WL}_$TAG.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$WL", d["config"]["kernel"], "value %.0f Msamp/s" % d["value"], "us/step %.2f" % (d["ms_per_step"]*1e3),
          "fp32 frac %.3f (peak %.1f TF)" % (r["frac"], r["peak"]), "hbm frac %.3f" % r["hbm"]["frac"],
          "e2e %.0f" % d["e2e"]["value"], "cpu", d["cpu_baseline"], "clocks", d["clocks"])
except Exception as e:
    print("$WL bench failed:", e); print(open("$OUT/bench_${WL}_$TAG.err").read()[-1500:])
PY
done
for TR in 2 4 8; do
  SPXB_TILED_TR=$TR timeout 300 python bench.py --workload C3 --steps 50 --warmup 5 --no-cpu-baseline --min-seconds 0.3 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('C3 TR=$TR us/step %.2f frac %.3f' % (d['ms_per_step']*1e3, d['roofline']['frac']))
"
done
# launch list (shares only) and one full capture of the tiled kernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 10 -c 120 --csv --log-file $OUT/launches_$TAG.csv \
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --min-seconds 0.001 > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tiled_fir -s 10 -c 2 -f -o $OUT/prof_tiled_$TAG \
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --min-seconds 0.001 > /dev/null 2>&1
ls -la $OUT | tail -20
