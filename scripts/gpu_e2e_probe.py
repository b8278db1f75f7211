#!/usr/bin/env python
"""Why is the C3 host-buffer pipeline at ~0.8 of the copy-only PCIe ceiling? Per-step period of
(a) free-running H2D + D2H on two streams, (b) D2H(k) chained behind H2D(k) by an event,
(c) a kernel on a third stream between them, (d) the library's submit/wait -- same bytes, same
rotating pinned buffers."""
import ctypes as C
import math
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import node_speex_resampler_b200 as pkg  # noqa: E402

S, ch, i, o, q, n = 1024, 2, 44100, 48000, 7, 882
cap = int(math.ceil(n * o / i))
in_bytes, out_bytes = S * n * ch * 2, S * cap * ch * 2
HR, STEPS = 6, 300
h_i = [torch.empty(in_bytes, dtype=torch.uint8).pin_memory() for _ in range(HR)]
h_o = [torch.empty(out_bytes, dtype=torch.uint8).pin_memory() for _ in range(HR)]
d_i = [torch.empty(in_bytes, dtype=torch.uint8, device="cuda") for _ in range(4)]
d_o = [torch.empty(out_bytes, dtype=torch.uint8, device="cuda") for _ in range(4)]
s_in, s_k, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, label):
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
    print(f"{label:58s} {(t1 - t0) / STEPS * 1e6:7.1f} us/step")


def free_running():
    for k in range(STEPS):
        with torch.cuda.stream(s_in):
            d_i[k % 4].copy_(h_i[k % HR], non_blocking=True)
        with torch.cuda.stream(s_out):
            h_o[k % HR].copy_(d_o[k % 4], non_blocking=True)


def chained(with_kernel):
    def run():
        for k in range(STEPS):
            with torch.cuda.stream(s_in):
                d_i[k % 4].copy_(h_i[k % HR], non_blocking=True)
                e1 = torch.cuda.Event()
                e1.record()
            if with_kernel:
                with torch.cuda.stream(s_k):
                    s_k.wait_event(e1)
                    d_o[k % 4][:in_bytes].copy_(d_i[k % 4], non_blocking=True)  # a ~4 us device kernel
                    e2 = torch.cuda.Event()
                    e2.record()
            else:
                e2 = e1
            with torch.cuda.stream(s_out):
                s_out.wait_event(e2)
                h_o[k % HR].copy_(d_o[k % 4], non_blocking=True)
    return run


def only(direction):
    def run():
        for k in range(STEPS):
            if direction == "h2d":
                with torch.cuda.stream(s_in):
                    d_i[k % 4].copy_(h_i[k % HR], non_blocking=True)
            else:
                with torch.cuda.stream(s_out):
                    h_o[k % HR].copy_(d_o[k % 4], non_blocking=True)
    return run


def lagged(lag):
    """D2H(k) waits for H2D(k + lag) -- negative lag = more slack than the real dependency"""
    def run():
        evs = []
        for k in range(STEPS + max(0, lag)):
            if k < STEPS or lag > 0:
                with torch.cuda.stream(s_in):
                    d_i[k % 4].copy_(h_i[k % HR], non_blocking=True)
                    e1 = torch.cuda.Event()
                    e1.record()
                    evs.append(e1)
            j = k - max(0, lag)
            if 0 <= j < STEPS:
                with torch.cuda.stream(s_out):
                    w = j + lag
                    if 0 <= w < len(evs):
                        s_out.wait_event(evs[w])
                    h_o[j % HR].copy_(d_o[j % 4], non_blocking=True)
    return run


def same_stream_d2h():
    """(g) D2H rides the compute stream right behind the kernel: one event per step"""
    for k in range(STEPS):
        with torch.cuda.stream(s_in):
            d_i[k % 4].copy_(h_i[k % HR], non_blocking=True)
            e1 = torch.cuda.Event()
            e1.record()
        with torch.cuda.stream(s_k):
            s_k.wait_event(e1)
            d_o[k % 4][:in_bytes].copy_(d_i[k % 4], non_blocking=True)
            h_o[k % HR].copy_(d_o[k % 4], non_blocking=True)


s_k2 = torch.cuda.Stream()


def two_compute_streams():
    """(h) kernels alternate between two compute streams (chained by a kernel-to-kernel event);
    each stream carries its own D2H behind its kernel"""
    prev = None
    for k in range(STEPS):
        with torch.cuda.stream(s_in):
            d_i[k % 4].copy_(h_i[k % HR], non_blocking=True)
            e1 = torch.cuda.Event()
            e1.record()
        sk = s_k if k % 2 == 0 else s_k2
        with torch.cuda.stream(sk):
            sk.wait_event(e1)
            if prev is not None:
                sk.wait_event(prev)
            d_o[k % 4][:in_bytes].copy_(d_i[k % 4], non_blocking=True)
            prev = torch.cuda.Event()
            prev.record()
            h_o[k % HR].copy_(d_o[k % 4], non_blocking=True)


timed(same_stream_d2h, "(g) H2D | kernel + D2H on one compute stream")
timed(two_compute_streams, "(h) H2D | kernel + D2H alternating on two compute streams")
timed(only("h2d"), "(f1) H2D only")
timed(only("d2h"), "(f2) D2H only")
timed(free_running, "(a) free-running H2D || D2H")
timed(lagged(-1), "(e1) D2H(k) behind H2D(k-1)")
timed(lagged(0), "(e2) D2H(k) behind H2D(k)")
timed(lagged(1), "(e3) D2H(k) behind H2D(k+1)")
timed(chained(False), "(b) D2H(k) behind H2D(k) (event)")
timed(chained(True), "(c) H2D -> kernel -> D2H (events, three streams)")

L = pkg.lib()
b = pkg.StreamBatch(S, ch, i, o, q)
hin = [L.spxb_host_alloc(in_bytes) for _ in range(HR)]
hout = [L.spxb_host_alloc(out_bytes) for _ in range(HR)]
nin, nout = np.empty(S, np.uint32), np.empty(S, np.uint32)
depth = L.spxb_batch_pipeline_depth(b._h)


def library():
    tickets = []
    for k in range(STEPS):
        nin.fill(n)
        nout.fill(cap)
        t = C.c_uint64(0)
        assert L.spxb_batch_submit(b._h, hin[k % HR], n, nin.ctypes.data, hout[k % HR], cap, nout.ctypes.data,
                                   C.byref(t)) == 0
        tickets.append(t.value)
        if k >= depth:
            L.spxb_batch_wait(b._h, tickets[k - depth])
    for t in tickets[-depth:]:
        L.spxb_batch_wait(b._h, t)


timed(library, f"(d) spxb_batch_submit/wait, depth {depth}")
