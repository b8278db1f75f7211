#!/usr/bin/env python
"""bench.py -- throughput of the Speex resampler hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload C3|C4|C5] [--kernel auto|strict|tiled|tensor]

One step = one 20 ms processChunk-equivalent for every stream of the batch (state carried
step to step). Workload (BASELINE.json): C3 = 1024 stereo streams 44100->48000 q7 per GPU
(the configuration the headline metric is quoted on), C4 = 4096 mono 48000->16000 q10,
C5 = 8192 stereo 96000->44100 q10 per GPU (65536 over 8). Streams are independent, so with
N GPUs every rank owns its own streams (weak scaling, no collective on the data path).

Prints ONE JSON line (rank 0):
  value      output Msamples/s, inputs already in HBM (hops queued in a ring > L2)
  e2e        same metric through the C ABI with pinned HOST buffers: H2D + kernel + D2H
             every step, pipelined with spxb_batch_submit / spxb_batch_wait
  roofline   of the FIR kernel. Tensor kernel (tcgen05 int8): the algorithmic work (2*N flops
             per output sample) takes less time at the measured tensor peak than
             the algorithmic bytes take at the measured HBM peak, so the binding roofline is
             HBM ("bound": "hbm"); the executed int8 MMA rate is reported beside it. FP32-FMA
             kernel ("tiled"): FP32 FMA rate against an FFMA probe measured in this run.
  cpu_baseline  the reference's own C (oracle/_ref, kind "reference") or the oracle port on
             this box's host cores, bounded sample of the same workload

--impl reference times that CPU implementation on the same config instead (rank 0 only).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (streams per GPU, channels, in_rate, out_rate, quality, frames in per 20 ms step)
    "C3": (1024, 2, 44100, 48000, 7, 882),
    "C4": (4096, 1, 48000, 16000, 10, 960),
    "C5": (8192, 2, 96000, 44100, 10, 1920),
    # not BASELINE shapes: mid-length filters used to place the lean / paced kernel threshold
    "X6": (1024, 2, 44100, 48000, 10, 882),   # N = 256: 6 stages
    "X8": (2048, 1, 32000, 16000, 10, 640),   # N = 512, mono, 2:1
}
L2_BYTES = 126 * 1024 * 1024


def describe(wl):
    S, ch, i, o, q, n = WORKLOADS[wl]
    return {"workload": f"{wl}: {S} streams/GPU x {ch} ch, {i}->{o} Hz, quality {q}, 20 ms steps "
                        f"({n} frames in), state carried", "streams_per_gpu": S, "channels": ch,
            "in_rate": i, "out_rate": o, "quality": q, "frames_in_per_step": n}


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, line in self.lines:
            if t0 is not None and not (t0 - 0.05 <= ts <= t1 + 0.1):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------
# CPU implementation of the path (reference arm / cpu_baseline)
# ---------------------------------------------------------------------------
CPU_WHAT = ("the reference's own deps/speex/resample.c compiled natively by oracle/Makefile "
            "(-O2 -ffp-contract=off -DFLOATING_POINT -DOUTSIDE_SPEEX): the native-C proxy for the shipped "
            "WASM module (no Node / WASM runtime in this image; the translated module itself runs at 0.9x of it)")


def cpu_run(wl: str, steps: int, warmup: int, threads: int | None = None):
    """The reference's CPU path on this box's cores: every step resamples one 20 ms chunk of
    every stream of the workload, streams spread over `threads` OS threads (ctypes releases
    the GIL inside the C call). Returns (samples/s, seconds per step, kind, threads)."""
    from oracle import oracle as O
    from node_speex_resampler_b200.signals import synth_pcm
    S, ch, i, o, q, n = WORKLOADS[wl]
    cls, kind = O.best_cpu_resampler()
    threads = threads or os.cpu_count() or 1
    threads = min(threads, S)
    cap = int(math.ceil(n * o / i)) + 1
    pcm = synth_pcm(min(S, 64), ch, n, i, seed=0xB200)
    streams = [cls(ch, i, o, q) for _ in range(S)]
    made = [0] * threads

    def worker(t, count):
        tot = 0
        for _ in range(count):
            for s in range(t, S, threads):
                y, _, m = streams[s].process(pcm[s % pcm.shape[0]], cap)
                tot += m * ch
        made[t] = tot

    def run(count):
        ts = [threading.Thread(target=worker, args=(t, count)) for t in range(threads)]
        t0 = time.perf_counter()
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        return time.perf_counter() - t0

    if warmup:
        run(warmup)
    dt = run(steps)
    return sum(made) / dt, dt / steps, kind, threads


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl = args.workload
    S, ch, i, o, q, n = WORKLOADS[wl]
    steps, W = args.steps, max(args.warmup, 3)
    # bound the run: ~0.75 core-seconds per C3 step; the warm-up is capped in TIME, not in the count
    # it reports (the CPU path has no clocks to ramp: two steps warm its caches and page in the streams)
    rate, sec_per_step, kind, threads = cpu_run(wl, steps, min(W, 2))
    sample = f"{steps} steps x {S} streams x 20 ms ({kind} C, one stream per thread, {threads} threads)"
    line = {"impl": "reference", "metric": "output_msamples_per_sec", "value": rate / 1e6, "unit": "Msamples/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": W, "ms_per_step": sec_per_step * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": describe(wl),
            "cpu_baseline": {"value": rate / 1e6, "unit": "Msamples/s", "cores": threads, "kind": kind,
                             "what": CPU_WHAT, "sample": sample, "warmup_steps_run": min(W, 2)},
            "e2e": {"value": rate / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------
class Ctx:
    pass


def measure(cx, wl, S, K, W, min_seconds, lean, kernel):
    """One workload on this rank's GPU: device-resident throughput (CUDA events), the roofline of
    its FIR kernel, and the same metric end to end through the C ABI with pinned host buffers.
    Returns a dict; the multi-rank reductions (max over ranks) happen inside."""
    torch, dist, pkg, _lib, L = cx.torch, cx.dist, cx.pkg, cx._lib, cx.L
    world, rank, local = cx.world, cx.rank, cx.local
    _, ch, i, o, q, n = WORKLOADS[wl]
    info = _lib.FilterInfo()
    L.spxb_filter_describe(i, o, q, C.byref(info))
    N = info.filt_len
    cap = int(math.ceil(n * o / i))
    batch = pkg.StreamBatch(S, ch, i, o, q, device=local)
    batch.set_kernel({"auto": pkg.KERNEL_AUTO, "strict": pkg.KERNEL_STRICT, "tiled": pkg.KERNEL_TILED,
                      "tensor": pkg.KERNEL_TENSOR}[kernel])
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)  # events below and the batch's kernels share this stream
    assert L.spxb_batch_set_stream(batch._h, C.c_void_p(stream.cuda_stream)) == 0

    # ---- inputs resident in HBM: a ring of distinct hops larger than L2 ----
    # rows padded to 16 bytes so the kernel's vector loads apply (unpadded rows also work,
    # through a slower frame-by-frame staging path)
    n_pad = (n * ch + 7) // 8 * 8 // ch
    cap_pad = (cap * ch + 7) // 8 * 8 // ch
    in_slot = S * n_pad * ch
    out_slot = S * cap_pad * ch
    ring = max(4, int(math.ceil(1.25 * L2_BYTES / (in_slot * 2))))
    ring = min(ring, 96)
    if cx.ring:
        ring = cx.ring
    # a ring that divides the step count makes every timed region walk the same slots, so the
    # library's CUDA graph of the hop sequence (spxb_batch_process_device_ring) is captured once
    for cand in range(ring, min(2 * ring, 96) + 1):
        if K % cand == 0 and not cx.ring:
            ring = cand
            break
    hop = pkg.synth_pcm(min(S, 256), ch, n * 4, i, seed=0xB200 + rank)
    d_in = torch.zeros((ring, S, n_pad * ch), dtype=torch.int16, device="cuda")
    for r in range(ring):
        sel = np.resize(np.arange(hop.shape[0]), S) if hop.shape[0] < S else np.arange(S)
        sl = hop[np.roll(sel, r), (r % 4) * n * ch:((r % 4) + 1) * n * ch]
        d_in[r, :, : n * ch].copy_(torch.from_numpy(np.ascontiguousarray(sl)))
    d_out = torch.zeros((ring, S, cap_pad * ch), dtype=torch.int16, device="cuda")

    def run_hops(first, count):
        e = L.spxb_batch_process_device_ring(batch._h, d_in.data_ptr(), n_pad, in_slot, d_out.data_ptr(), cap_pad,
                                             out_slot, ring, n, cap, first, count)
        if e:
            raise RuntimeError(_lib.strerror(e) + ": " + _lib.last_error())

    run_hops(0, W)
    cx.barrier()
    # probe one step to size the repetitions (each timed region is EXACTLY K steps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run_hops(W, K)
    e1.record()
    torch.cuda.synchronize()
    est = max(e0.elapsed_time(e1) * 1e-3, 1e-6)
    reps = int(min(400, max(3, math.ceil(min_seconds / est))))
    if world > 1:
        # every repetition is bracketed by barriers: the ranks must agree on how many there are (each
        # rank's own estimate can differ by one, which left one rank alone in a collective for ten minutes)
        t = torch.tensor([reps], device="cuda", dtype=torch.int64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        reps = int(t.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.15)
    c0 = batch.counters().kernel_launches
    times = []
    cx.barrier()
    t_start = time.perf_counter()
    step_no = W + K
    for r in range(reps):
        cx.barrier()
        e0.record()
        run_hops(step_no, K)
        e1.record()
        cx.barrier()
        times.append(e0.elapsed_time(e1) * 1e-3)
        step_no += K
    t_end = time.perf_counter()
    launches = (batch.counters().kernel_launches - c0) // reps
    clocks = sampler.stop(t_start, t_end) if rank == 0 else None
    sec = statistics.median(times)
    if world > 1:
        t = torch.tensor([sec], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
    out_samples_step = S * cap * ch  # per GPU
    value = out_samples_step * world * K / sec
    kernel_used = {0: "auto", 1: "strict", 2: "tiled", 3: "tensor"}[batch.last_kernel()]

    # ---- roofline of the FIR kernel ----
    # --lean: no auxiliary kernels (FFMA peak probe, device-copy floor), so that an ncu launch list
    # of the command holds the step's own kernels only
    fp32_peak = 148 * 128 * 2 * 1.965e9 if lean else cx.fp32_peak()
    flops_per_launch = out_samples_step * 2.0 * N
    t_launch = sec / K
    bytes_per_launch = (out_samples_step * 2.0 + S * n * ch * 2.0 + 2.0 * S * ch * (N - 1) * 2.0)
    peaks = cx.peaks
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    # DRAM bytes per launch of this kernel from the committed ncu --set full capture (never measured
    # under the profiler here; profiles/traffic.json is written by scripts/ncu_traffic.py from the
    # .ncu-rep files); null when no capture exists for the workload at this batch size
    traffic, traffic_src = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(f"{wl}:{kernel_used}")
        if tr and int(tr.get("streams", WORKLOADS[wl][0])) == S:
            traffic = tr["dram_read_bytes"] + tr["dram_write_bytes"]
            traffic_src = tr["source"]
    except Exception:
        pass
    hbm = {"achieved": bytes_per_launch / t_launch / 1e9, "peak": hbm_peak, "unit": "GB/s",
           "frac": bytes_per_launch / t_launch / 1e9 / hbm_peak,
           "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650",
           "bytes_per_output_sample": bytes_per_launch / out_samples_step}
    fma = {"achieved": flops_per_launch / t_launch / 1e12, "peak": fp32_peak / 1e12, "unit": "TFLOP/s",
           "frac": (flops_per_launch / t_launch) / fp32_peak if fp32_peak else None,
           "peak_source": "FFMA probe measured in this run (MEASURED_PEAKS.json has no fp32 entry); "
                          "nominal 148 SM x 128 lanes x 2 x 1.965 GHz = 74.4",
           "flops_per_output_sample": 2 * N}
    # What a launch that only MOVES the algorithmic bytes costs in the same loop: torch's device copy
    # kernel over bytes/2 in + bytes/2 out, K back to back on the same stream. At C3's 8.6 MB this
    # is dominated by per-launch latency, not by HBM -- a practical floor for one step.
    copy_floor = None
    try:
        if lean:
            raise RuntimeError("skipped (--lean)")
        half = int(bytes_per_launch // 2) // 16 * 16
        c_src = torch.empty((8, half), dtype=torch.uint8, device="cuda")
        c_dst = torch.empty_like(c_src)
        for timed in (False, True):
            e0.record()
            for k in range(K):
                c_dst[k % 8].copy_(c_src[k % 8], non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
        copy_floor = {"us_per_launch": e0.elapsed_time(e1) * 1e3 / K, "bytes_moved": 2 * half,
                      "how": f"torch copy_ (at::direct_copy_kernel) of {half} B device to device -- read + write "
                             f"= the step's algorithmic bytes -- {K} launches back to back on the stream, CUDA "
                             "events; includes per-launch latency, like the kernel's own number"}
        del c_src, c_dst
    except Exception as ex:  # noqa: BLE001
        copy_floor = {"error": str(ex)}
    if kernel_used == "tensor":
        # binding roofline: max(algorithmic bytes / HBM peak, algorithmic flops / tensor peak)
        bf16_peak = float(peaks.get("bf16_tflops", 1590.0)) * 1e12
        t_hbm = bytes_per_launch / (hbm_peak * 1e9)
        t_tensor = flops_per_launch / bf16_peak
        geom = batch.tensor_geometry() or {}
        mma_ops = (geom.get("tiles", 0) * geom.get("groups", 0) * geom.get("ksteps", 0)
                   * 2 * 128 * 3 * geom.get("nt", 0) * 32 * 2.0)
        roof = dict(hbm, bound="hbm" if t_hbm >= t_tensor else "tensor", traffic=traffic, traffic_source=traffic_src,
                    launch_us=t_launch * 1e6, algorithmic_bytes_per_launch=bytes_per_launch,
                    why_bound=f"algorithmic bytes / measured HBM peak = {t_hbm * 1e6:.2f} us vs algorithmic "
                              f"2N flops / measured bf16 tensor peak = {t_tensor * 1e6:.2f} us per launch",
                    tensor={"executed_int8_ops_per_launch": mma_ops, "achieved": mma_ops / t_launch / 1e12,
                            "unit": "Tops/s (int8 MMA, incl. band padding and the 6 digit products)",
                            "peak_nominal": 4500.0, "frac_of_nominal": mma_ops / t_launch / 4.5e15,
                            "geometry": geom},
                    fp32_fma_equivalent=fma, d2d_memcpy_same_bytes=copy_floor)
    else:
        roof = dict(fma, bound="fp32_fma", traffic=traffic, traffic_source=traffic_src, launch_us=t_launch * 1e6,
                    hbm=hbm, d2d_memcpy_same_bytes=copy_floor)

    # ---- end to end through the C ABI with pinned host buffers ----
    L.spxb_batch_use_own_stream(batch._h)
    hr = min(6, ring)
    in_bytes, out_bytes = S * n * ch * 2, S * cap * ch * 2
    hin = [L.spxb_host_alloc(in_bytes) for _ in range(hr)]
    hout = [L.spxb_host_alloc(out_bytes) for _ in range(hr)]
    src = np.ascontiguousarray(d_in[:hr, :, : n * ch].cpu().numpy())
    for k in range(hr):
        C.memmove(hin[k], src[k].ctypes.data, in_bytes)
    del src
    nin = np.empty(S, np.uint32)
    nout = np.empty(S, np.uint32)
    depth = L.spxb_batch_pipeline_depth(batch._h)

    host_t = {"submit": 0.0, "wait": 0.0}

    def e2e_steps(count):
        tickets = []
        nin_p, nout_p = nin.ctypes.data, nout.ctypes.data
        for k in range(count):
            ta = time.perf_counter()
            nin.fill(n)
            nout.fill(cap)
            t = C.c_uint64(0)
            e = L.spxb_batch_submit(batch._h, hin[k % hr], n, nin_p, hout[k % hr], cap, nout_p, C.byref(t))
            if e:
                raise RuntimeError(_lib.strerror(e) + ": " + _lib.last_error())
            tickets.append(t.value)
            tb = time.perf_counter()
            if k >= depth:
                L.spxb_batch_wait(batch._h, tickets[k - depth])
            tc = time.perf_counter()
            host_t["submit"] += tb - ta
            host_t["wait"] += tc - tb
        for t in tickets[-depth:]:
            L.spxb_batch_wait(batch._h, t)

    Ke = max(K, 20)
    e2e_steps(max(W, 3))
    batch.synchronize()
    host_t["submit"] = host_t["wait"] = 0.0
    cx.barrier()
    e0.record()
    tw0 = time.perf_counter()
    e2e_steps(Ke)
    batch.synchronize()
    e1.record()
    cx.barrier()
    tw1 = time.perf_counter()
    e2e_sec = max(e0.elapsed_time(e1) * 1e-3, tw1 - tw0)
    if world > 1:
        t = torch.tensor([e2e_sec], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_sec = float(t.item())
    e2e_value = out_samples_step * world * Ke / e2e_sec
    # the ceiling of that number: this step's H2D and D2H bytes copied concurrently from/to pinned
    # memory with nothing else running on this GPU (every rank does it at the same time, so at N > 1
    # it is the ceiling UNDER the contention for the host's memory and PCIe switches), no kernel
    pcie = None
    try:
        # same number of distinct pinned host buffers as the e2e loop rotates over (a single
        # re-used buffer stays in the host's last-level cache and overstates the ceiling)
        h_i = [torch.empty(in_bytes, dtype=torch.uint8).pin_memory() for _ in range(hr)]
        h_o = [torch.empty(out_bytes, dtype=torch.uint8).pin_memory() for _ in range(hr)]
        d_i = torch.empty(in_bytes, dtype=torch.uint8, device="cuda")
        d_o = torch.empty(out_bytes, dtype=torch.uint8, device="cuda")
        s_a, s_b = torch.cuda.Stream(), torch.cuda.Stream()
        nrep = max(10, int(0.2 / max(e2e_sec / Ke, 1e-6)))
        for timed in (False, True):
            torch.cuda.synchronize()
            cx.barrier()
            tp0 = time.perf_counter()
            for k in range(nrep):
                with torch.cuda.stream(s_a):
                    d_i.copy_(h_i[k % hr], non_blocking=True)
                with torch.cuda.stream(s_b):
                    h_o[k % hr].copy_(d_o, non_blocking=True)
            torch.cuda.synchronize()
            tp1 = time.perf_counter()
        per_step = (tp1 - tp0) / nrep
        if world > 1:
            t = torch.tensor([per_step], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            per_step = float(t.item())
        pcie = {"copy_only_us_per_step": per_step * 1e6, "h2d_GBs": in_bytes / per_step / 1e9,
                "d2h_GBs": out_bytes / per_step / 1e9, "host_buffers": hr,
                "ceiling_msamples_per_sec": out_samples_step * world / per_step / 1e6,
                "e2e_frac_of_ceiling": (e2e_value / (out_samples_step * world / per_step)),
                "how": "this rank's H2D and D2H of one step's bytes, concurrently, all ranks at once, max over ranks"}
        del h_i, h_o, d_i, d_o
    except Exception as ex:  # noqa: BLE001
        pcie = {"error": str(ex)}
    for p in hin + hout:
        L.spxb_host_free(p)
    batch.close()
    del d_in, d_out
    torch.cuda.empty_cache()
    return {"config": dict(describe(wl), streams_per_gpu=S,
                           workload=describe(wl)["workload"].replace(f"{WORKLOADS[wl][0]} streams/GPU", f"{S} streams/GPU")),
            "value": value / 1e6, "unit": "Msamples/s", "ms_per_step": sec / K * 1e3,
            "measurement": {"timing": f"median of {reps} regions of exactly {K} steps, CUDA events; inputs: ring of "
                                      f"{ring} distinct hops in HBM ({ring * in_slot * 2 >> 20} MiB > L2), no L2 flush needed",
                            "kernel": kernel_used, "filt_len": N},
            "dtype": "s8 x s8 -> s32 (exact integer FIR)" if kernel_used == "tensor" else "f32",
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": e2e_value / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": in_bytes,
                    "d2h_bytes_per_step": out_bytes, "steps": Ke, "pipeline_depth": depth,
                    "how": "spxb_batch_submit/wait (C ABI), pinned host buffers, H2D+kernel+D2H per step",
                    "host_us_per_step": {"submit": host_t["submit"] / Ke * 1e6, "wait": host_t["wait"] / Ke * 1e6},
                    "pcie": pcie},
            "roofline": roof}


def single_stream_latency(cx):
    """The reference's real call pattern (src/index.ts:96-102): ONE stereo stream, one 20 ms chunk per
    synchronous call through speex_resampler_process_interleaved_int with pageable host buffers.
    Wall time per call, median of 300, per kernel family, beside the reference's C on one core."""
    pkg, _lib, L = cx.pkg, cx._lib, cx.L
    from oracle import oracle as O
    ch, i, o, q, n = 2, 44100, 48000, 7, 882
    cap = 960
    x = pkg.synth_pcm(1, ch, n * 8, i, seed=77)[0]
    out = np.zeros(cap * ch, np.int16)
    res = {}
    for name, k in (("strict", pkg.KERNEL_STRICT), ("tensor", pkg.KERNEL_TENSOR), ("auto", pkg.KERNEL_AUTO)):
        err = C.c_int(0)
        st = L.speex_resampler_init(ch, i, o, q, C.byref(err))
        L.spxb_batch_set_kernel(L.spxb_resampler_batch(st), k)
        ts = []
        for it in range(340):
            chunk = x[(it % 8) * n * ch:((it % 8) + 1) * n * ch]
            n_in, n_out = C.c_uint32(n), C.c_uint32(cap)
            t0 = time.perf_counter()
            e = L.speex_resampler_process_interleaved_int(st, chunk.ctypes.data, C.byref(n_in), out.ctypes.data, C.byref(n_out))
            ts.append(time.perf_counter() - t0)
            assert e == 0 and n_out.value in (959, 960)
        L.speex_resampler_destroy(st)
        res[name] = statistics.median(ts[40:]) * 1e6
    cls, kind = O.best_cpu_resampler()
    r = cls(ch, i, o, q)
    ts = []
    for it in range(340):
        chunk = x[(it % 8) * n * ch:((it % 8) + 1) * n * ch]
        t0 = time.perf_counter()
        r.process(chunk, cap)
        ts.append(time.perf_counter() - t0)
    res["cpu_" + kind] = statistics.median(ts[40:]) * 1e6
    res["what"] = ("one stereo 44100->48000 q7 stream, 882 frames per call, speex_resampler_process_interleaved_int, "
                   "pageable buffers, wall clock around the call (includes both copies and the synchronisation), "
                   "median of 300 calls; strict is the default of the drop-in surface")
    return res


def file_configs(cx):
    """BASELINE configs 1-2: the reference's own resources/*.pcm through processChunk -- the whole file
    in one call (src/test.ts:24-44, what its test prints the time of) and in 20 ms chunks -- beside the
    reference's C on one core, same file, same calls. Wall clock; the GPU arm includes the copies."""
    pkg = cx.pkg
    from oracle import oracle as O
    cases = [("44100hz_test.pcm", 2, 44100, 48000, 7), ("24000hz_mono_test.pcm", 1, 24000, 44100, 1),
             ("24000hz_mono_test.pcm", 1, 24000, 44100, 7), ("24000hz_mono_test.pcm", 1, 24000, 44100, 10)]
    if O.fixture_path(cases[0][0]) is None:
        return {"unavailable": "oracle/_ref/resources absent (make -C oracle where /root/reference exists)"}
    cls, kind = O.best_cpu_resampler()
    rows = []
    for f, ch, i, o, q in cases:
        data = open(O.fixture_path(f), "rb").read()
        hop = i // 50 * ch * 2  # 20 ms of input, in bytes
        row = {"file": f, "channels": ch, "in_rate": i, "out_rate": o, "quality": q, "bytes": len(data)}
        for kname, kern in (("strict", pkg.KERNEL_STRICT), ("tensor", pkg.KERNEL_TENSOR)):
            r = pkg.SpeexResampler(ch, i, o, q)
            r.kernel = kern
            r.processChunk(data[: hop * 4])  # init + first-call planning outside the timed region
            r.destroy()
            r = pkg.SpeexResampler(ch, i, o, q)
            r.kernel = kern
            r.processChunk(b"")
            t0 = time.perf_counter()
            y = r.processChunk(data)
            t1 = time.perf_counter()
            r.destroy()
            r = pkg.SpeexResampler(ch, i, o, q)
            r.kernel = kern
            r.processChunk(b"")
            t2 = time.perf_counter()
            total = 0
            for k in range(0, len(data) - len(data) % (ch * 2), hop):
                total += len(r.processChunk(data[k:k + hop]))
            t3 = time.perf_counter()
            r.destroy()
            row[kname] = {"one_shot_ms": (t1 - t0) * 1e3, "one_shot_msamples_per_sec": len(y) / 2 / (t1 - t0) / 1e6,
                          "chunked_20ms_ms": (t3 - t2) * 1e3, "chunked_us_per_call": (t3 - t2) / max(1, -(-len(data) // hop)) * 1e6}
        c = cls(ch, i, o, q)
        t0 = time.perf_counter()
        yc = c.processChunk(data)
        t1 = time.perf_counter()
        c = cls(ch, i, o, q)
        t2 = time.perf_counter()
        for k in range(0, len(data) - len(data) % (ch * 2), hop):
            c.processChunk(data[k:k + hop])
        t3 = time.perf_counter()
        row["cpu_" + kind] = {"one_shot_ms": (t1 - t0) * 1e3, "one_shot_msamples_per_sec": len(yc) / 2 / (t1 - t0) / 1e6,
                              "chunked_20ms_ms": (t3 - t2) * 1e3, "cores": 1}
        row["out_frames"] = len(yc) // 2 // ch
        rows.append(row)
    return {"what": "whole file in one processChunk call and in 20 ms chunks, wall clock, one stream; "
                    "cpu = " + CPU_WHAT, "cases": rows}


def ours(args):
    import torch
    import torch.distributed as dist

    import node_speex_resampler_b200 as pkg
    from node_speex_resampler_b200 import _lib

    cx = Ctx()
    cx.torch, cx.dist, cx.pkg, cx._lib = torch, dist, pkg, _lib
    cx.world = world = int(os.environ.get("WORLD_SIZE", "1"))
    cx.rank = rank = int(os.environ.get("RANK", "0"))
    cx.local = local = int(os.environ.get("LOCAL_RANK", "0"))
    cx.ring = args.ring
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path "
                         "(use --impl reference for the CPU implementation)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    cx.barrier = barrier
    cx.L = L = pkg.lib()
    try:
        cx.peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        cx.peaks = {}
    _fp32 = []

    def fp32_peak():
        if not _fp32:
            _fp32.append(L.spxb_measure_fp32_peak(8192))
        return _fp32[0]
    cx.fp32_peak = fp32_peak

    wl = args.workload
    S = args.streams or WORKLOADS[wl][0]
    # weak scaling: the job is world * S independent streams, rank r owns one contiguous block
    from node_speex_resampler_b200.sharding import shard_range
    lo, hi = shard_range(S * world, world, rank)
    assert hi - lo == S
    K, W = args.steps, max(args.warmup, 3)
    head = measure(cx, wl, S, K, W, args.min_seconds, args.lean, args.kernel)

    # ---- the other BASELINE shapes, timed like the headline (BASELINE configs[3], configs[4]) ----
    also = None
    if wl == "C3" and not args.no_also and not args.streams:
        also = {}
        # config 5 as BASELINE states it: 65536 stereo streams sharded over the GPUs of the run
        # (2/4/8); on one GPU the per-GPU share of the 8-GPU case
        shapes = {"C4": WORKLOADS["C4"][0], "C5": 8192 if world == 1 else 65536 // world}
        for name, s_gpu in shapes.items():
            try:
                m = measure(cx, name, s_gpu, K, W, min(args.min_seconds, 0.4), True, args.kernel)
                m["scaling"] = "weak" if name == "C4" else ("strong (65536 streams over the run's GPUs)" if world > 1 else
                                                            "weak (8192 streams: one GPU's share of the 8-GPU case)")
                r = m["roofline"]
                m["roofline"] = {k: r.get(k) for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "traffic_source",
                                                       "launch_us", "algorithmic_bytes_per_launch", "bytes_per_output_sample")}
                also[name] = m
            except Exception as ex:  # noqa: BLE001
                also[name] = {"error": repr(ex)}

    # ---- single stream: call latency and the reference's own files (rank 0, one GPU) ----
    latency = files = None
    if rank == 0 and world == 1 and not args.no_also and not args.streams:
        try:
            latency = single_stream_latency(cx)
        except Exception as ex:  # noqa: BLE001
            latency = {"error": repr(ex)}
        try:
            files = file_configs(cx)
        except Exception as ex:  # noqa: BLE001
            files = {"error": repr(ex)}

    # ---- CPU baseline beside it (rank 0, N == 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        S_, ch_, i_, o_, q_, n_ = WORKLOADS[wl]
        per_step_core_s = {"C3": 0.75, "C4": 0.9, "C5": 30.0}.get(wl, 2.0)
        cores = os.cpu_count() or 1
        cpu_steps = max(1, int(20.0 * cores / per_step_core_s / cores)) if wl != "C5" else 1
        cpu_steps = max(1, min(cpu_steps, int(15.0 * cores / per_step_core_s)))
        rate, sps, kind, threads = cpu_run(wl, cpu_steps, 0)
        cpu = {"value": rate / 1e6, "unit": "Msamples/s", "cores": threads, "kind": kind, "what": CPU_WHAT,
               "sample": f"{cpu_steps} steps x {S_} streams x 20 ms, one stream per thread, {threads} threads"}

    if rank == 0:
        line = {"metric": "output_msamples_per_sec", "value": head["value"], "unit": "Msamples/s", "n_gpus": world,
                "steps": K, "warmup": W, "ms_per_step": head["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": head["dtype"], "data": "synthetic",
                "config": describe(wl) if not args.streams else head["config"],
                "measurement": head["measurement"],
                "clocks": head["clocks"], "gpu_launches": head["gpu_launches"],
                "e2e": head["e2e"], "roofline": head["roofline"], "cpu_baseline": cpu,
                "also": also, "latency_us": latency, "files": files}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3", choices=sorted(WORKLOADS))
    ap.add_argument("--kernel", default="auto", choices=["auto", "strict", "tiled", "tensor"])
    ap.add_argument("--min-seconds", type=float, default=1.0, help="clock-sampling window for the timed regions")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-also", action="store_true",
                    help="headline workload only: skip the `also` (C4, C5), `latency_us` and `files` blocks")
    ap.add_argument("--streams", type=int, default=0, help="experiments: streams per GPU instead of the workload's")
    ap.add_argument("--ring", type=int, default=0, help="experiments: number of distinct hops resident in HBM")
    ap.add_argument("--lean", action="store_true", help="skip the auxiliary probes (FFMA peak, device-copy floor)")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
