#!/usr/bin/env python
"""Many consecutive calls of one batch through the host-buffer entry (debugging aid): prints after every call."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import node_speex_resampler_b200 as pkg  # noqa: E402

S, ch, i, o, q, n = 8192, 2, 96000, 44100, 10, 1920
calls = int(sys.argv[1]) if len(sys.argv) > 1 else 30
cap = -(-n * o // i)
b = pkg.StreamBatch(S, ch, i, o, q)
b.set_kernel(pkg.KERNEL_TENSOR)
pcm = pkg.synth_pcm(64, ch, n, i, seed=3)
pcm = np.ascontiguousarray(np.resize(pcm, (S, n * ch)))
for k in range(calls):
    out, used, made = b.process(pcm, n, cap)
    print(k, int(used[0]), int(made[0]), int(np.abs(out[0].astype(np.int64)).sum()), b.tensor_geometry(), flush=True)
