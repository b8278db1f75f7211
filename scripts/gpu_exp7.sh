#!/usr/bin/env bash
# one-visit experiment: does the persistent kernel's PCM load throughput follow the L1 left over by
# its shared memory? (x-stage cap -> smaller dynamic shared memory -> larger L1 carve-out)
set -u
cd "$(dirname "$0")/.."
run() { local label=$1; shift
  env "$@" 2>/dev/null | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); g = d['roofline']['tensor']['geometry']; print('$label us/step %.2f  nt %d stages %d smem %d tiles %d' % (d['ms_per_step']*1e3, g['nt'], g['stages'], g['smem_bytes'], g['tiles']))
"
}
B="timeout 200 python bench.py --kernel tensor --steps 192 --warmup 10 --no-cpu-baseline --no-also --lean --min-seconds 0.3"
for XS in 2 3 4 6; do
  run "v3 C5 packed nt64 xstages<=$XS" SPXB_UMMA_RESIDENT=1 SPXB_UMMA_NT=64 SPXB_UMMA2_XSTAGES=$XS $B --workload C5
  run "v3 C5 packed nt48 xstages<=$XS" SPXB_UMMA_RESIDENT=1 SPXB_UMMA_NT=48 SPXB_UMMA2_XSTAGES=$XS $B --workload C5
  run "v3 C5 dense nt64 xstages<=$XS" SPXB_UMMA_RESIDENT=1 SPXB_UMMA_DENSE=1 SPXB_UMMA_NT=64 SPXB_UMMA2_XSTAGES=$XS $B --workload C5
  run "v3 C5 packed nt112 xstages<=$XS" SPXB_UMMA_RESIDENT=1 SPXB_UMMA2_XSTAGES=$XS $B --workload C5
  run "v3 C3 dense nt64 xstages<=$XS" SPXB_UMMA_RESIDENT=1 SPXB_UMMA_DENSE=1 SPXB_UMMA_NT=64 SPXB_UMMA2_XSTAGES=$XS $B --workload C3
  run "v3 C3 packed nt112 xstages<=$XS" SPXB_UMMA_RESIDENT=1 SPXB_UMMA2_XSTAGES=$XS $B --workload C3
done
