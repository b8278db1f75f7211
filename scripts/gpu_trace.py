#!/usr/bin/env python
"""Timeline of the tensor kernel's CTAs (SPXB_UMMA_TRACE=1): where a tile's time goes.
usage: SPXB_UMMA_TRACE=1 python scripts/gpu_trace.py [C3 C4 C5]"""
import ctypes as C
import os
import sys

import numpy as np

os.environ.setdefault("SPXB_UMMA_TRACE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import node_speex_resampler_b200 as pkg  # noqa: E402

SHAPES = {"C3": (1024, 2, 44100, 48000, 7, 882), "C4": (4096, 1, 48000, 16000, 10, 960),
          "C5": (8192, 2, 96000, 44100, 10, 1920)}
NAMES = {0: "start", 1: "setup done", 2: "fetch0 issued", 3: "stage0 stored", 4: "convert loop end",
         5: "acc ready (conv)", 6: "epilogue math done", 7: "outputs stored", 8: "history done", 9: "exit",
         30: "convert loop end (warp 7)", 11: "mma: all issued", 12: "tma: first bulk", 13: "tma: last bulk"}
for it in range(12):
    NAMES[20 + it] = f"mma: stage {it} full"
L = pkg.lib()
for wl in (sys.argv[1:] or ["C3", "C5"]):
    S, ch, i, o, q, n = SHAPES[wl]
    S = int(os.environ.get("STREAMS", S))
    cap = -(-n * o // i)
    b = pkg.StreamBatch(S, ch, i, o, q)
    b.set_kernel(pkg.KERNEL_TENSOR)
    pcm = pkg.synth_pcm(min(S, 64), ch, n, i, seed=3)
    pcm = np.ascontiguousarray(np.resize(pcm, (S, n * ch)))
    for k in range(4):
        b.process(pcm, n, cap)
    geom = b.tensor_geometry()
    ctas = geom["tiles"] * (geom["groups"] + 3)
    buf = np.zeros(ctas * 32, np.uint64)
    got = L.spxb_batch_tensor_trace(b._h, buf.ctypes.data, buf.size)
    t = buf.reshape(-1, 32)[:got].astype(np.int64)
    print(f"== {wl}: geom {geom}, {got} CTAs traced; SM clock cycles after CTA start (median | p10 | p90 over CTAs)")
    rel = t - t[:, :1]
    for slot in sorted(NAMES):
        col = rel[:, slot]
        col = col[t[:, slot] != 0] if slot else col
        if col.size == 0:
            continue
        print(f"  {NAMES[slot]:>22s}: {np.median(col):9.0f} | {np.percentile(col, 10):9.0f} | {np.percentile(col, 90):9.0f}")
    gt0, gt1 = t[:, 14], t[:, 15]
    print(f"  globaltimer: first start -> last end {int(gt1.max() - gt0.min())} ns; CTA duration median "
          f"{np.median(gt1 - gt0):.0f} ns; start spread {int(gt0.max() - gt0.min())} ns; "
          f"distinct SMs {len(set(t[:, 16].tolist()))}")
    # waves: CTAs per SM
    b.close()
