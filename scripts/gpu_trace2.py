#!/usr/bin/env python
"""Timeline of the PERSISTENT tensor kernel's CTAs (kernels_umma2.cu built with -DSPXB_UMMA2_TRACE:
`make -C node_speex_resampler_b200/csrc trace`): where a CTA's and a tile's time goes.
usage: SPXB_LIB_PATH=$PWD/node_speex_resampler_b200/libspeexb200_trace.so python scripts/gpu_trace2.py [C3 C4 C5]"""
import os
import sys

import numpy as np

os.environ.setdefault("SPXB_UMMA_TRACE", "1")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import node_speex_resampler_b200 as pkg  # noqa: E402

SHAPES = {"C3": (1024, 2, 44100, 48000, 7, 882), "C4": (4096, 1, 48000, 16000, 10, 960),
          "C5": (8192, 2, 96000, 44100, 10, 1920), "X6": (1024, 2, 44100, 48000, 10, 882)}
NAMES = {0: "start", 5: "mma warp: before barrier init", 6: "mma warp: barriers initialised", 7: "loader: past griddepcontrol.wait", 1: "past the first __syncthreads", 2: "first fetches issued", 3: "before griddepcontrol.wait", 4: "after griddepcontrol.wait", 12: "tap loads issued", 8: "history done",
         11: "mma: all issued", 10: "converters done", 9: "exit"}
TILE = ["", "epi: waits for the accumulator", "", "epi: tile stored",
        "mma: accumulator set free", "mma: stage 0 full", "mma: last stage issued"]
L = pkg.lib()
for wl in (sys.argv[1:] or ["C3", "C4", "C5"]):
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
    buf = np.zeros(160 * 128, np.uint64)
    got = L.spxb_batch_tensor_trace(b._h, buf.ctypes.data, buf.size)
    t = buf.reshape(-1, 128)[:got].astype(np.int64)
    print(f"== {wl}: geom {geom}, {got} CTAs traced; SM clock cycles after CTA start (median | p10 | p90 over CTAs)")
    rel = t - t[:, :1]

    def show(name, slot):
        col = rel[:, slot][t[:, slot] != 0] if slot else rel[:, slot]
        if col.size:
            print(f"  {name:>28s}: {np.median(col):9.0f} | {np.percentile(col, 10):9.0f} | {np.percentile(col, 90):9.0f}   ({col.size} CTAs)")
    for slot in (0, 5, 6, 1, 7, 12, 3, 4, 2):
        show(NAMES[slot], slot)
    for j in range(5):
        for k, nm in enumerate(TILE):
            if nm:
                show(f"tile {j} {nm}", 16 + 8 * j + k)
    for k, nm in enumerate(("step: enter", "step: slot empty", "step: converted + stored", "step: fenced + arrived", "step: next loads issued")):
        show(f"converter step n_iters+3 {nm}", 56 + k)
    for it in range(12):
        for k, nm in enumerate(("wait", "full", "issued")):
            show(f"mma tile 1 stage {it} {nm}", 64 + 3 * it + k)
    for it in range(12):
        show(f"tile 1 stage {it}: box issued", 112 + it)
        show(f"tile 1 stage {it}: raw bytes seen by its converter group", 100 + it)
    for slot in (8, 10, 11, 9):
        show(NAMES[slot], slot)
    gt0, gt1 = t[:, 14], t[:, 15]
    print(f"  globaltimer: first start -> last end {int(gt1.max() - gt0.min())} ns; CTA duration median "
          f"{np.median(gt1 - gt0):.0f} ns; start spread {int(gt0.max() - gt0.min())} ns; distinct SMs {len(set(t[:, 13].tolist()))}")
    b.close()
