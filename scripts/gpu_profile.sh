#!/usr/bin/env bash
# Round profile bundle: full bench lines, ncu launch list, ncu --set full of the FIR kernel.
set -u
TAG=${1:-r1}
OUT=gpurun_out; mkdir -p $OUT
for WL in C3 C4 C5; do
  timeout 900 python bench.py --workload $WL > $OUT/bench_${WL}_$TAG.json 2> $OUT/bench_${WL}_$TAG.err
  tail -c 400 $OUT/bench_${WL}_$TAG.json; echo
done
timeout 600 python bench.py --impl reference --steps 20 --warmup 2 > $OUT/bench_ref_C3_$TAG.json 2>/dev/null
# every launch of a short run with its device time (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 5 -c 200 --csv --log-file $OUT/launches_$TAG.csv \
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --min-seconds 0.001 > /dev/null 2>&1
for WL in C3 C5; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:stream_fir -s 8 -c 1 -f -o $OUT/prof_${WL}_$TAG \
    python bench.py --workload $WL --steps 10 --warmup 3 --no-cpu-baseline --min-seconds 0.001 > /dev/null 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:strict_fir -s 2 -c 1 -f -o $OUT/prof_strict_C3_$TAG \
  python bench.py --workload C3 --kernel strict --steps 5 --warmup 3 --no-cpu-baseline --min-seconds 0.001 > /dev/null 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/clocks_$TAG.csv
ls -la $OUT | tail -15
