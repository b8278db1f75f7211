#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -5 | tee $OUT/pytest_graph.log
run() {
  local label=$1; shift
  for WL in C3 C4 C5; do
    env "$@" timeout 300 python bench.py --workload $WL --steps 200 --warmup 10 --no-cpu-baseline --min-seconds 0.5 2>$OUT/err.txt | python -c "
import sys, json
ok=False
for l in sys.stdin:
    if l.startswith('{'):
        ok=True
        d = json.loads(l)
        print('$label $WL us/step %.2f launches %d e2e %.0f (%.2f of pcie)' % (d['ms_per_step']*1e3, d['gpu_launches'], d['e2e']['value'], d['e2e']['pcie']['e2e_frac_of_ceiling']))
if not ok: print('$label $WL FAILED', open('$OUT/err.txt').read()[-800:])
"
  done
}
(run graph X=1; run nograph SPXB_RING_GRAPH=0; SPXB_RING_GRAPH=1 timeout 300 python bench.py --workload C3 --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -c 400) | tee $OUT/graph_ab.log
