#!/usr/bin/env python
"""Single-stream call latency by chunk size and kernel family (the reference's real call pattern:
one SpeexResampler per stream, synchronous processChunk). Wall clock around
speex_resampler_process_interleaved_int with pageable buffers, median of 200 calls.
usage: python scripts/gpu_latency_table.py"""
import ctypes as C
import os
import statistics
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import node_speex_resampler_b200 as pkg  # noqa: E402
from oracle import oracle as O  # noqa: E402

L = pkg.lib()
cls, kind = O.best_cpu_resampler()
print(f"{'config':28s} {'frames':>7s} {'strict us':>10s} {'tensor us':>10s} {'cpu us':>9s}  (cpu = {kind} C, one core)")
for ch, i, o, q in ((2, 44100, 48000, 7), (1, 48000, 16000, 10), (2, 96000, 44100, 10)):
    for n in (160, i // 100, i // 50, i // 10, i):
        cap = -(-n * o // i) + 1
        x = pkg.synth_pcm(1, ch, n * 4, i, seed=9)[0]
        out = np.zeros(cap * ch, np.int16)
        row = {}
        for name, k in (("strict", pkg.KERNEL_STRICT), ("tensor", pkg.KERNEL_TENSOR)):
            err = C.c_int(0)
            st = L.speex_resampler_init(ch, i, o, q, C.byref(err))
            L.spxb_batch_set_kernel(L.spxb_resampler_batch(st), k)
            ts = []
            for it in range(230):
                chunk = x[(it % 4) * n * ch:((it % 4) + 1) * n * ch]
                n_in, n_out = C.c_uint32(n), C.c_uint32(cap)
                t0 = time.perf_counter()
                e = L.speex_resampler_process_interleaved_int(st, chunk.ctypes.data, C.byref(n_in), out.ctypes.data, C.byref(n_out))
                ts.append(time.perf_counter() - t0)
                assert e == 0
            L.speex_resampler_destroy(st)
            row[name] = statistics.median(ts[30:]) * 1e6
        r = cls(ch, i, o, q)
        ts = []
        for it in range(60):
            chunk = x[(it % 4) * n * ch:((it % 4) + 1) * n * ch]
            t0 = time.perf_counter()
            r.process(chunk, cap)
            ts.append(time.perf_counter() - t0)
        row["cpu"] = statistics.median(ts[10:]) * 1e6
        print(f"{ch}ch {i}->{o} q{q:<2d}         {n:7d} {row['strict']:10.1f} {row['tensor']:10.1f} {row['cpu']:9.1f}")
