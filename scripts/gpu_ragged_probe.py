#!/usr/bin/env python
"""What cohort grouping buys a ragged batch: 1024 stereo streams 44.1k -> 48k q7 in two cohorts of
512 at different stream positions, 20 ms hops from device buffers; AUTO (tensor launches over the
cohorts' id lists) against the strict kernel over the whole batch."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import node_speex_resampler_b200 as pkg  # noqa: E402

L = pkg.lib()
S, ch, i, o, q, n, cap = 1024, 2, 44100, 48000, 7, 882, 962
for cohorts in (2, 4, 8):
    for name, kernel in (("auto (cohorts on the tensor kernel)", pkg.KERNEL_AUTO), ("strict", pkg.KERNEL_STRICT)):
        b = pkg.StreamBatch(S, ch, i, o, q)
        b.set_kernel(kernel)
        ts = torch.cuda.Stream()
        torch.cuda.set_stream(ts)
        assert L.spxb_batch_set_stream(b._h, C.c_void_p(ts.cuda_stream)) == 0
        x = torch.randint(-8000, 8000, (S, n * ch), dtype=torch.int16, device="cuda")
        y = torch.zeros((S, cap * ch), dtype=torch.int16, device="cuda")
        # stagger the cohorts: cohort c first eats 100*c + 37 frames on its own
        which = np.arange(S) * cohorts // S
        for c in range(1, cohorts):
            nin = np.where(which == c, 100 * c + 37, 0).astype(np.uint32)
            nout = np.where(which == c, cap, 0).astype(np.uint32)
            assert L.spxb_batch_process_device(b._h, x.data_ptr(), n, nin.ctypes.data, y.data_ptr(), cap,
                                               nout.ctypes.data) == 0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        K = 50
        for rep in range(3):
            e0.record()
            for k in range(K):
                nin = np.full(S, n, np.uint32)
                nout = np.full(S, cap, np.uint32)
                assert L.spxb_batch_process_device(b._h, x.data_ptr(), n, nin.ctypes.data, y.data_ptr(), cap,
                                                   nout.ctypes.data) == 0, pkg._lib.last_error()
            e1.record()
            torch.cuda.synchronize()
        lk = b.last_kernel()
        print(f"{cohorts} cohorts, {name:38s}: {e0.elapsed_time(e1) * 1e3 / K:8.1f} us per hop (last kernel family {lk}, "
              f"{(b.counters().kernel_launches) } launches total)")
        b.close()
