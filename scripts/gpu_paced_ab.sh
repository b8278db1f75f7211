set -u
for ROUND in 1 2; do
  for P in auto 1 0; do
    for WL in C3 C4 C5; do
      if [ $P = auto ]; then unset SPXB_UMMA_PACED; else export SPXB_UMMA_PACED=$P; fi
      SPXB_LIB_PATH=$PWD/ab/lib_new.so timeout 300 python bench.py --workload $WL --kernel tensor --steps 200 --warmup 10 --no-cpu-baseline --lean --min-seconds 0.3 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('paced=$P round $ROUND $WL us/step %.2f' % (d['ms_per_step']*1e3))
"
    done
  done
done
unset SPXB_UMMA_PACED
SPXB_LIB_PATH=$PWD/ab/lib_trace_2end.so timeout 300 python bench.py --workload C5 --kernel tensor --steps 200 --warmup 10 --no-cpu-baseline --lean --min-seconds 0.3 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print('trace_2end C5 us/step %.2f' % (d['ms_per_step']*1e3))
"
timeout 600 python -m pytest tests -m gpu -q -x --timeout 600 2>&1 | tail -3
