#!/usr/bin/env python
"""Host<->device copy bandwidth on this box (pinned memory): the ceiling of the end-to-end number."""
import json
import sys
import time

import torch

res = {}
for mb in (3.6, 7.9, 63, 256):
    n = int(mb * 1e6)
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    reps = max(5, int(2e9 / n))

    def run(h2d, d2h):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if h2d:
                with torch.cuda.stream(s1):
                    d_in.copy_(h_in, non_blocking=True)
            if d2h:
                with torch.cuda.stream(s2):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        return n * reps / (time.perf_counter() - t0) / 1e9

    run(True, True)
    res[f"{mb}MB"] = {"h2d_GBs": round(run(True, False), 2), "d2h_GBs": round(run(False, True), 2),
                      "both_each_GBs": round(run(True, True), 2)}
print(json.dumps(res))
