#!/usr/bin/env bash
# Tensor-kernel bring-up visit: probe, diagnostics, tests, benches per kernel.
set -u
TAG=${1:-t2}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv | tee $OUT/smi_$TAG.txt
make -s -C node_speex_resampler_b200/csrc probe 2>&1 | tail -3
timeout 120 node_speex_resampler_b200/csrc/umma_probe 2>&1 | tail -40 | tee $OUT/probe_$TAG.log
echo "probe rc=$?"
timeout 300 python scripts/gpu_tensor_check.py 2>&1 | tail -40 | tee $OUT/tensor_check_$TAG.log
echo "tensor_check rc=${PIPESTATUS[0]}"
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -25 | tee $OUT/pytest_$TAG.log
for WL in C3 C4 C5; do
 for K in tiled tensor; do
  timeout 300 python bench.py --workload $WL --kernel $K --steps 50 --warmup 5 --no-cpu-baseline --min-seconds 0.5 > $OUT/bench_${WL}_${K}_$TAG.json 2> $OUT/bench_${WL}_${K}_$TAG.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_${WL}_${K}_$TAG.json").read().strip().splitlines()[-1])
    r = d["roofline"]
    print("$WL $K", d["config"]["kernel"], "value %.0f Msamp/s" % d["value"], "us/step %.2f" % (d["ms_per_step"]*1e3),
          "frac %.3f (peak %.1f %s)" % (r["frac"], r["peak"], r["bound"]), "e2e %.0f" % d["e2e"]["value"], "clocks", d["clocks"])
except Exception as e:
    print("$WL $K bench failed:", e); print(open("$OUT/bench_${WL}_${K}_$TAG.err").read()[-1500:])
PY
 done
done
