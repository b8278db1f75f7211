#!/usr/bin/env bash
set -u
OUT=gpurun_out; mkdir -p $OUT
(timeout 300 python scripts/gpu_graph_probe.py C3 C4 C5
 echo "--- PDL off"
 SPXB_UMMA_PDL=0 timeout 300 python scripts/gpu_graph_probe.py C3 C4) 2>&1 | tee $OUT/graph_probe.log
timeout 300 python bench.py --workload C3 --steps 200 --warmup 5 --no-cpu-baseline --min-seconds 0.3 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); e = d['e2e']
        print('C3 e2e %.0f (%.2f of pcie; ceiling copy-only %.1f us/step) d2d copy floor %s' % (e['value'], e['pcie']['e2e_frac_of_ceiling'], e['pcie']['copy_only_us_per_step'], d['roofline']['d2d_memcpy_same_bytes']))
" | tee -a $OUT/graph_probe.log
