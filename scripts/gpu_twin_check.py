#!/usr/bin/env python
"""Two batches, the same input (the second one permuted): where do their outputs differ? (debugging aid;
the loop of tests/test_parity_gpu.py::test_full_size_long_filter_properties, repeated, with a report)
usage: gpu_twin_check.py S ch in_rate out_rate quality frames [calls [rounds]]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import node_speex_resampler_b200 as pkg  # noqa: E402

S, ch, i, o, q, n = map(int, sys.argv[1:7])
calls = int(sys.argv[7]) if len(sys.argv) > 7 else 3
rounds = int(sys.argv[8]) if len(sys.argv) > 8 else 1
cap = -(-n * o // i)
perm = np.random.default_rng(11).permutation(S)
base = pkg.synth_pcm(256, ch, n * calls, i, seed=0xC0FFEE)
sel = np.resize(np.arange(256), S)
total_bad = 0
for rnd in range(rounds):
    b, twin = pkg.StreamBatch(S, ch, i, o, q), pkg.StreamBatch(S, ch, i, o, q)
    for k in range(calls):
        time.sleep(0.05 * (rnd % 3))
        pcm = np.ascontiguousarray(base[np.roll(sel, 3 * k), k * n * ch:(k + 1) * n * ch])
        pcm[7] = pcm[5]
        out, used, made = b.process(pcm, n, cap)
        out2, _, _ = twin.process(pcm[perm], n, cap)
        ref = out[perm]
        bad = np.argwhere(out2 != ref)
        if len(bad) == 0:
            continue
        total_bad += 1
        print(f"round {rnd} call {k}: kernels {b.last_kernel()} {twin.last_kernel()} geom {b.tensor_geometry()} "
              f"mismatching samples {len(bad)}")
        rows = np.unique(bad[:, 0])
        print("  twin rows", rows[:20], "... =", len(rows), "rows; original streams", perm[rows[:20]])
        r0 = rows[0]
        cols = bad[bad[:, 0] == r0][:, 1]
        print("  first row: columns", cols[:24], "...", len(cols), "of", out.shape[1], "| values", out2[r0, cols[:6]], ref[r0, cols[:6]])
        print("  groups of the twin rows:", np.unique(rows // (128 // ch))[:20], " lanes:", np.unique(rows % (128 // ch))[:40])
        print("  groups of the original rows:", np.unique(perm[rows] // (128 // ch))[:20])
        d = np.abs(out2.astype(np.int64) - ref.astype(np.int64))
        print("  max |difference|", int(d.max()), " columns touched overall:", np.unique(bad[:, 1])[:40])
    b.close()
    twin.close()
print(f"{rounds} rounds x {calls} calls: {total_bad} calls with mismatches")
