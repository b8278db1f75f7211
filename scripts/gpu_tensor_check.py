#!/usr/bin/env python
"""Bring-up diagnostics for the tensor kernel on a B200: small batches against the oracle,
printing where (stream, channel, output index) mismatches fall instead of just failing."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import node_speex_resampler_b200 as pkg  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = [
    ("C3", 70, 2, 44100, 48000, 7, 882, 3),
    ("C4", 130, 1, 48000, 16000, 10, 960, 3),
    ("C5", 33, 2, 96000, 44100, 10, 1920, 2),
    ("up2", 37, 1, 24000, 48000, 5, 480, 3),
    ("odd", 5, 2, 44100, 48000, 3, 1000, 4),
]
bad_total = 0
for name, S, ch, i, o, q, n, calls in CASES:
    b = pkg.StreamBatch(S, ch, i, o, q)
    b.set_kernel(pkg.KERNEL_TENSOR)
    refs = [O.OracleResampler(ch, i, o, q) for _ in range(S)]
    cap = int(np.ceil(n * o / i)) + 2
    for k in range(calls):
        pcm = pkg.synth_pcm(S, ch, n, i, seed=0xC0DE, start_frame=k * n)
        try:
            out, used, made = b.process(pcm, n, cap)
        except Exception as e:  # noqa: BLE001
            print(name, "call", k, "FAILED:", e, pkg._lib.last_error())
            bad_total += 1
            break
        geom = b.tensor_geometry()
        worst, nbad, nbig, total = 0, 0, 0, 0
        first = None
        for s in range(S):
            y, u, m = refs[s].process(pcm[s], cap)
            if (u, m) != (int(used[s]), int(made[s])):
                print(name, "lengths differ", s, (u, m), (int(used[s]), int(made[s])))
                bad_total += 1
            d = np.abs(y.astype(np.int32) - out[s, : m * ch].astype(np.int32))
            total += d.size
            nbad += int((d != 0).sum())
            nbig += int((d > 1).sum())
            if d.max(initial=0) > 1 and first is None:
                idx = int(np.argmax(d > 1))
                first = (s, idx // ch, idx % ch, int(y[idx]), int(out[s, idx]))
            worst = max(worst, int(d.max(initial=0)))
        print(f"{name} call {k}: kernel={b.last_kernel()} geom={geom} max|d|={worst} "
              f"off-by-one={nbad - nbig}/{total} wrong={nbig} first_wrong(stream,frame,ch,want,got)={first}", flush=True)
        bad_total += nbig
    for s in (0, S - 1):
        ls, fr, mg, hist = b.get_state(s)
        rls, rfr, rhist = refs[s].state(0)
        ok = (ls, fr) == (rls, rfr) and np.array_equal(hist.reshape(-1, ch)[:, 0].astype(np.float32), rhist)
        if not ok:
            print(name, "state mismatch stream", s, (ls, fr), (rls, rfr))
            bad_total += 1
    b.close()
print("tensor check:", "OK" if bad_total == 0 else f"{bad_total} problems")
sys.exit(0 if bad_total == 0 else 1)
