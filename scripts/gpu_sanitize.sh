#!/usr/bin/env bash
# compute-sanitizer over the three kernel families on small shapes (sanitizers)
set -u
OUT=gpurun_out; mkdir -p $OUT
cat > /tmp/san_case.py <<'PY'
import numpy as np, sys
sys.path.insert(0, ".")
import node_speex_resampler_b200 as pkg
from oracle import oracle as O
for kernel in (pkg.KERNEL_TENSOR, pkg.KERNEL_TILED, pkg.KERNEL_STRICT):
    for (S, ch, i, o, q, n) in ((70, 2, 44100, 48000, 7, 882), (130, 1, 48000, 16000, 10, 960)):
        b = pkg.StreamBatch(S, ch, i, o, q); b.set_kernel(kernel)
        cap = int(np.ceil(n * o / i)) + 1
        r = O.OracleResampler(ch, i, o, q)
        for k in range(2):
            pcm = pkg.synth_pcm(S, ch, n, i, seed=5, start_frame=k * n)
            out, used, made = b.process(pcm, n, cap)
            y, u, m = r.process(pcm[0], cap)
            assert np.abs(y.astype(int) - out[0, : m * ch].astype(int)).max() <= 1
        b.close()
# ragged cohorts + float batch
b = pkg.StreamBatch(80, 2, 44100, 48000, 5)
n_in = np.where(np.arange(80) < 40, 500, 333).astype(np.uint32)
b.process(pkg.synth_pcm(80, 2, 500, 44100, seed=1), n_in, 600)
b.process(pkg.synth_pcm(80, 2, 500, 44100, seed=2), n_in, 600)
b.close()
f = pkg.StreamBatch(3, 2, 44100, 48000, 7, sample_format="f32")
f.process_f32(np.random.default_rng(0).standard_normal((3, 1000)).astype(np.float32), 500, 600)
f.close()
print("sanitizer case ran to the end")
PY
for TOOL in memcheck racecheck; do
  echo "== compute-sanitizer --tool $TOOL"
  timeout 600 compute-sanitizer --tool $TOOL --print-limit 20 python /tmp/san_case.py 2>&1 | grep -v "^$" | tail -15
done | tee $OUT/sanitizer.log
