#!/usr/bin/env bash
# Round-1 evidence bundle: GPU tests, full bench lines (ours + reference arm), ncu launch list,
# ncu --set full of the tensor kernel (C3, C5), clocks. Outputs in gpurun_out/ (copied to profiles/ by hand).
set -u
TAG=${1:-r1b}
OUT=gpurun_out; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -6 | tee $OUT/pytest_$TAG.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $OUT/smoke_$TAG.log
for WL in C3 C4 C5; do
  timeout 900 python bench.py --workload $WL > $OUT/bench_${WL}_$TAG.json 2> $OUT/bench_${WL}_$TAG.err
  tail -c 300 $OUT/bench_${WL}_$TAG.json; echo
done
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/bench_ref_C3_$TAG.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 5 -c 200 --csv --log-file $OUT/launches_$TAG.csv \
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline --min-seconds 0.001 > /dev/null 2>&1
for WL in C3 C5; do
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:umma_fir -s 8 -c 1 -f -o $OUT/prof_umma_${WL}_$TAG \
    python bench.py --workload $WL --kernel tensor --steps 10 --warmup 3 --no-cpu-baseline --min-seconds 0.001 > /dev/null 2>&1
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $OUT/clocks_$TAG.csv
ls -la $OUT | tail -12
