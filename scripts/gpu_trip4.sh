#!/usr/bin/env bash
# quick iteration visit: parity diagnostics, timeline, per-kernel bench lines
set -u
TAG=${1:-t4}
OUT=gpurun_out
mkdir -p $OUT
timeout 300 python scripts/gpu_tensor_check.py 2>&1 | tail -22 | tee $OUT/tensor_check_$TAG.log
SPXB_UMMA_TRACE=1 timeout 300 python scripts/gpu_trace.py C3 C4 C5 2>&1 | tee $OUT/trace_$TAG.log
for WL in C3 C4 C5; do
  timeout 300 python bench.py --workload $WL --kernel tensor --steps 50 --warmup 5 --no-cpu-baseline --min-seconds 0.5 > $OUT/bench_${WL}_tensor_$TAG.json 2> $OUT/bench_${WL}_tensor_$TAG.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_${WL}_tensor_$TAG.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$WL", d["config"]["kernel"], "value %.0f Msamp/s" % d["value"], "us/step %.2f" % (d["ms_per_step"]*1e3),
          "frac %.3f (peak %.1f %s)" % (r["frac"], r["peak"], r["bound"]), "e2e %.0f" % d["e2e"]["value"], r.get("tensor", {}).get("geometry"))
except Exception as e:
    print("$WL bench failed:", e); print(open("$OUT/bench_${WL}_tensor_$TAG.err").read()[-1500:])
PY
done
if [ "${2:-}" = "tests" ]; then timeout 1200 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -8 | tee $OUT/pytest_$TAG.log; fi
