set -u
for ROUND in 1 2; do
  for P in 0 1; do
    for WL in X6 X8; do
      SPXB_UMMA_PACED=$P timeout 300 python bench.py --workload $WL --kernel tensor --steps 200 --warmup 10 --no-cpu-baseline --lean --min-seconds 0.3 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('paced=$P round $ROUND $WL us/step %.2f' % (d['ms_per_step']*1e3), d['roofline']['tensor']['geometry'])
"
    done
  done
done
