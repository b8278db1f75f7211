#!/usr/bin/env bash
# Evidence bundle (current build): GPU tests, smoke, full bench lines (ours + reference arm), ncu
# launch list, ncu --set full captures of the tensor kernel (C3, C5), pipeline-slot sweep, clocks.
# Outputs land in gpurun_out/; summaries are copied into profiles/ by hand
# (scripts/ncu_summary.py, scripts/launch_summary.py).
set -u
TAG=${1:-r1c}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -6 | tee $OUT/pytest_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke_$TAG.log
# the default line (C3 headline + also{C4, C5} + latency_us + files + cpu_baseline), as the driver runs it
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > $OUT/bench_$TAG.json 2> $OUT/bench_$TAG.err
tail -c 300 $OUT/bench_$TAG.json; echo
timeout 600 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $OUT/bench_ref_C3_$TAG.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 5 -c 200 --csv --log-file $OUT/launches_$TAG.csv \
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-also --lean --min-seconds 0.001 > /dev/null 2>&1
for WL in C3 C4 C5; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:umma[23]?_fir -s 8 -c 1 -f -o $OUT/prof_umma_${WL}_$TAG \
    python bench.py --workload $WL --kernel tensor --steps 10 --warmup 3 --no-cpu-baseline --no-also --min-seconds 0.001 > /dev/null 2>&1
done
for SL in 2 3 4; do
  SPXB_PIPELINE_SLOTS=$SL timeout 300 python bench.py --workload C3 --steps 200 --warmup 5 --no-cpu-baseline --no-also --min-seconds 0.3 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); e = d['e2e']
        print('slots $SL C3 e2e %.0f (%.2f of pcie) host us/step %s' % (e['value'], e['pcie']['e2e_frac_of_ceiling'], e['host_us_per_step']))
"
done | tee $OUT/e2e_slots_$TAG.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/clocks_$TAG.csv
ls -la $OUT | tail -12
