#!/usr/bin/env python
"""Is the C3-sized step launch-bound? The same K hops timed (a) launched one by one on a stream,
(b) captured once into a CUDA graph and replayed. usage: python scripts/gpu_graph_probe.py [C3 C4 C5]"""
import ctypes as C
import math
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import node_speex_resampler_b200 as pkg  # noqa: E402

SHAPES = {"C3": (1024, 2, 44100, 48000, 7, 882), "C4": (4096, 1, 48000, 16000, 10, 960),
          "C5": (8192, 2, 96000, 44100, 10, 1920)}
L = pkg.lib()
K = int(os.environ.get("K", 200))
for wl in (sys.argv[1:] or ["C3"]):
    S, ch, i, o, q, n = SHAPES[wl]
    cap = int(math.ceil(n * o / i))
    b = pkg.StreamBatch(S, ch, i, o, q)
    b.set_kernel(pkg.KERNEL_TENSOR)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert L.spxb_batch_set_stream(b._h, C.c_void_p(stream.cuda_stream)) == 0
    n_pad = (n * ch + 7) // 8 * 8 // ch
    cap_pad = (cap * ch + 7) // 8 * 8 // ch
    in_slot, out_slot = S * n_pad * ch, S * cap_pad * ch
    ring = min(96, max(4, int(math.ceil(1.25 * 126e6 / (in_slot * 2)))))
    ring = K // max(1, K // ring)  # ring divides K: every replay walks the same slots
    while K % ring:
        ring += 1
    d_in = torch.randint(-8000, 8000, (ring, S, n_pad * ch), dtype=torch.int16, device="cuda")
    d_out = torch.zeros((ring, S, cap_pad * ch), dtype=torch.int16, device="cuda")

    def hops(first, count):
        e = L.spxb_batch_process_device_ring(b._h, d_in.data_ptr(), n_pad, in_slot, d_out.data_ptr(), cap_pad,
                                             out_slot, ring, n, cap, first, count)
        assert e == 0, (e, pkg._lib.last_error())

    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    hops(0, 2 * K)
    torch.cuda.synchronize()
    ts = []
    for r in range(20):
        e0.record()
        hops(0, K)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3 / K)
    print(f"{wl} stream launches: {np.median(ts):.2f} us/step (min {min(ts):.2f}) ring {ring}")
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            hops(0, K)
        torch.cuda.synchronize()
        g.replay()
        torch.cuda.synchronize()
        ts = []
        for r in range(20):
            e0.record()
            g.replay()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3 / K)
        print(f"{wl} graph replay:    {np.median(ts):.2f} us/step (min {min(ts):.2f})")
    except Exception as ex:  # noqa: BLE001
        print(f"{wl} graph capture failed: {ex}")
    b.close()
