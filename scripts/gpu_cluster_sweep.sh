#!/usr/bin/env bash
# cluster-size sweep of the tensor kernel (tap tiles multicast across the CTAs of a cluster)
set -u
OUT=gpurun_out; mkdir -p $OUT
for CL in 1 2 4; do
  for WL in C3 C4 C5; do
    SPXB_UMMA_CLUSTER=$CL timeout 300 python bench.py --workload $WL --kernel tensor --steps 200 --warmup 10 --no-cpu-baseline --no-also --min-seconds 0.3 2>$OUT/err.txt | python -c "
import sys, json
ok = False
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); ok = True
        print('cluster $CL $WL us/step %.2f' % (d['ms_per_step']*1e3), d['roofline']['tensor']['geometry'])
if not ok: print('cluster $CL $WL FAILED', open('$OUT/err.txt').read()[-600:])
"
  done
done | tee $OUT/cluster_sweep.log
