#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q --timeout 600 2>&1 | tail -15 | tee $OUT/pytest_f32.log
for WL in C3 C4 C5; do
  timeout 300 python bench.py --workload $WL --steps 200 --warmup 10 --no-cpu-baseline --min-seconds 0.5 2>$OUT/err.txt | python -c "
import sys, json
ok=False
for l in sys.stdin:
    if l.startswith('{'):
        ok=True
        d = json.loads(l)
        print('graph $WL us/step %.2f launches %d e2e %.0f (%.2f of pcie)' % (d['ms_per_step']*1e3, d['gpu_launches'], d['e2e']['value'], d['e2e']['pcie']['e2e_frac_of_ceiling']))
if not ok: print('$WL FAILED', open('$OUT/err.txt').read()[-800:])
"
done | tee $OUT/graph_ab2.log
