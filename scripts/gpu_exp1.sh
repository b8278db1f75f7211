#!/usr/bin/env bash
# one-visit experiments: L2 prefetch distance A/B, L2-resident input vs HBM ring, v1 vs v2
set -u
P=node_speex_resampler_b200
run() { # label, env..., then bench args
  local label=$1; shift
  env "$@" 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('$label us/step %.2f' % (d['ms_per_step']*1e3))
"
}
B="timeout 200 python bench.py --kernel tensor --steps 192 --warmup 10 --no-cpu-baseline --no-also --lean --min-seconds 0.3"
for WL in C5 C4 C3; do
  run "v1 $WL" SPXB_UMMA_RESIDENT=0 $B --workload $WL
  run "v2 $WL" $B --workload $WL
  run "v2-pf4 $WL" SPXB_LIB_PATH=$PWD/$P/lib_pf4.so $B --workload $WL
  run "v2-pf8 $WL" SPXB_LIB_PATH=$PWD/$P/lib_pf8.so $B --workload $WL
done
for S in 2048; do
  run "v2 C5 streams=$S ring=default" $B --workload C5 --streams $S
  run "v2 C5 streams=$S ring=1 (L2 resident)" $B --workload C5 --streams $S --ring 1
  run "v1 C5 streams=$S ring=default" SPXB_UMMA_RESIDENT=0 $B --workload C5 --streams $S
  run "v1 C5 streams=$S ring=1 (L2 resident)" SPXB_UMMA_RESIDENT=0 $B --workload C5 --streams $S --ring 1
done
