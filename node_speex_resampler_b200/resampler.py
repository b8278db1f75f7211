"""Host-side mirror of the reference's TypeScript surface (src/index.ts) over the C ABI.

Same names, argument meaning and error behaviour as the reference:

* ``SpeexResampler.initPromise``                          src/index.ts:31
* ``SpeexResampler(channels, inRate, outRate, quality=7)``  src/index.ts:40-44
* ``SpeexResampler.processChunk(chunk) -> bytes``         src/index.ts:50-116
* ``SpeexResamplerTransform`` (``_transform``)            src/index.ts:121-162
* ``SpeexResampler.processChunks(resamplers, chunks)``    new: one launch for many streams

Node is not available in this image, so the mirror is Python (ctypes) instead of
TypeScript + N-API; INTEGRATION.md shows the N-API binding a maintainer would add. All
arithmetic runs in libspeexb200.so on the GPU; nothing here touches sample values.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Callable, Iterable, List, Optional, Sequence

import numpy as np

from . import _lib
from ._lib import KERNEL_AUTO, KERNEL_STRICT, KERNEL_TENSOR, KERNEL_TILED  # noqa: F401


class _Resolved:
    """Stand-in for the module-load promise of src/index.ts:19/31: loading is synchronous
    here, so the promise is already settled; it can be awaited or ``.result()``-ed."""

    def __init__(self, loader: Callable[[], object]):
        self._loader = loader

    def result(self):
        return self._loader()

    def __await__(self):
        async def _coro():
            return self._loader()
        return _coro().__await__()

    def then(self, fn):
        return fn(self._loader())


def _load_module():
    L = _lib.lib()
    if L.spxb_device_count() <= 0:
        raise RuntimeError("no CUDA device: node_speex_resampler_b200 has no CPU path")
    return L


def _chunk_bytes(chunk) -> bytes:
    if isinstance(chunk, np.ndarray):
        return chunk.tobytes()
    return bytes(chunk)


class SpeexResampler:
    """Drop-in for the reference class of the same name (src/index.ts:21-117)."""

    initPromise = _Resolved(_load_module)
    # Kernel family (not in the reference). The drop-in surface -- processChunk and the Transform
    # built on it -- defaults to the STRICT kernel, which reproduces the reference's bytes exactly;
    # the tensor-core kernel (+-1 LSB, >= 90 dB, the north star's tolerance) is opt-in there
    # (``r.kernel = KERNEL_TENSOR`` or KERNEL_AUTO). The batched entry the north star adds,
    # processChunks, defaults to AUTO (tensor kernel whenever the call qualifies) because
    # throughput is its purpose; set ``batch_kernel = KERNEL_STRICT`` for reference bytes there too.
    kernel = KERNEL_STRICT
    batch_kernel = KERNEL_AUTO

    def __init__(self, channels, inRate, outRate, quality=7):
        # like the reference constructor: no validation, no side effects (index.ts:40-44)
        self.channels = channels
        self.inRate = inRate
        self.outRate = outRate
        self.quality = quality
        self._resamplerPtr = None
        self._outBufferSize = -1
        self._group: Optional["StreamBatch"] = None
        self._group_index = -1

    # -- lifetime --------------------------------------------------------------
    def _ensure_init(self):
        if self._resamplerPtr:
            return
        L = _lib.lib()
        err = C.c_int(0)
        ptr = L.speex_resampler_init(int(self.channels), int(self.inRate), int(self.outRate),
                                     int(self.quality), C.byref(err))
        if err.value != 0 or not ptr:
            # index.ts:61-65: throw strerror text; _resamplerPtr stays falsy so a retry re-inits
            detail = _lib.last_error()
            msg = _lib.strerror(err.value)
            raise RuntimeError(msg + (f" ({detail})" if detail and err.value == 1 else ""))
        self._resamplerPtr = ptr
        L.spxb_batch_set_kernel(L.spxb_resampler_batch(ptr), int(self.kernel))

    def destroy(self):
        """Not in the reference (its instances leak); frees the device state."""
        if self._resamplerPtr:
            _lib.lib().speex_resampler_destroy(self._resamplerPtr)
            self._resamplerPtr = None

    def __del__(self):
        try:
            self.destroy()
        except Exception:
            pass

    # -- the capacity rule of index.ts:80-95 -----------------------------------
    def _capacity_frames(self, nbytes: int) -> int:
        target = math.ceil(nbytes * self.outRate / self.inRate)
        if self._outBufferSize < target:
            self._outBufferSize = target
        # setValue(ptr, size / channels / 2, 'i32') truncates toward zero
        return int(self._outBufferSize / self.channels / 2)

    def processChunk(self, chunk) -> bytes:
        """Resample one chunk of interleaved int16 PCM (index.ts:50-116)."""
        if self._group is not None:
            raise RuntimeError("this stream is attached to a batch; use processChunks")
        data = _chunk_bytes(chunk)
        frame_bytes = self.channels * 2
        # JS: `length % 0` is NaN, and NaN !== 0, so zero channels trips this same check
        if frame_bytes == 0 or len(data) % frame_bytes != 0:
            raise ValueError("Chunk length should be a multiple of channels * 2 bytes")
        self._ensure_init()
        L = _lib.lib()
        cap = self._capacity_frames(len(data))
        n_in = C.c_uint32(len(data) // self.channels // 2)
        n_out = C.c_uint32(cap)
        out = np.empty(max(cap * self.channels, 1), dtype=np.int16)
        src = np.frombuffer(data, dtype=np.int16) if data else np.zeros(1, np.int16)
        err = L.speex_resampler_process_interleaved_int(
            self._resamplerPtr, src.ctypes.data, C.byref(n_in), out.ctypes.data, C.byref(n_out))
        if err != 0:
            raise RuntimeError(_lib.strerror(err))
        # index.ts:108-115: the consumed count is ignored, the result is a fresh copy
        return out[: n_out.value * self.channels].tobytes()

    # -- batched entry (north star: processChunks) -------------------------------
    @staticmethod
    def processChunks(resamplers: Sequence["SpeexResampler"], chunks: Sequence) -> List[bytes]:
        """Resample chunk i with resampler i in as few launches as possible. The result equals
        ``[r.processChunk(c) for r, c in zip(resamplers, chunks)]``. Streams that share
        (channels, inRate, outRate, quality) form one device batch (one launch); the first call
        binds them to it, later calls must pass the same resamplers in the same order."""
        if len(resamplers) != len(chunks):
            raise ValueError("processChunks needs one chunk per resampler")
        if not resamplers:
            return []
        # group by configuration, keeping the caller's order inside each group
        by_config = {}
        for idx, r in enumerate(resamplers):
            by_config.setdefault((r.channels, r.inRate, r.outRate, r.quality), []).append(idx)
        if len(by_config) == 1:
            members = resamplers
            idx_lists = [None]
        else:
            idx_lists = list(by_config.values())
        out: List[bytes] = [b""] * len(resamplers)
        for idx in idx_lists:
            members = resamplers if idx is None else [resamplers[k] for k in idx]
            group = members[0]._group
            # same resamplers, same order -- compared element by element, by identity (a caller may
            # reorder its own list in place between calls)
            if group is None or len(group.members) != len(members) or \
                    any(a is not b for a, b in zip(group.members, members)):
                group = StreamBatch._adopt(members)
            res = group.processChunks(chunks if idx is None else [chunks[k] for k in idx])
            if idx is None:
                return res
            for k, y in zip(idx, res):
                out[k] = y
        return out


class StreamBatch:
    """n_streams streams of one (channels, inRate, outRate, quality) on one GPU, processed
    together. Backed by spxb_batch_* (include/speexb200.h part 2)."""

    def __init__(self, n_streams, channels, inRate, outRate, quality=7, device=0, sample_format="s16"):
        L = _lib.lib()
        err = C.c_int(0)
        if sample_format not in ("s16", "f32"):
            raise ValueError("sample_format is 's16' or 'f32'")
        # 'f32': float history (the reference's own `mem` type); serves process_f32 -- the Speex
        # float entry -- and int16 calls on one state, bit-exactly, on the strict kernel
        self.sample_format = sample_format
        create = L.spxb_batch_create_f32 if sample_format == "f32" else L.spxb_batch_create
        self._h = create(int(n_streams), int(channels), int(inRate), int(outRate),
                         int(quality), int(device), C.byref(err))
        if not self._h:
            detail = _lib.last_error()
            raise RuntimeError(_lib.strerror(err.value) + (f" ({detail})" if detail else ""))
        self.n_streams, self.channels = int(n_streams), int(channels)
        self.inRate, self.outRate, self.quality, self.device = inRate, outRate, quality, device
        self._out_buffer_size = np.full(self.n_streams, -1.0)
        self.members: Sequence[SpeexResampler] = ()

    @classmethod
    def _adopt(cls, resamplers: Sequence[SpeexResampler]) -> "StreamBatch":
        r0 = resamplers[0]
        key = (r0.channels, r0.inRate, r0.outRate, r0.quality)
        for r in resamplers:
            if (r.channels, r.inRate, r.outRate, r.quality) != key:
                raise ValueError("processChunks needs resamplers of one (channels, rates, quality)")
            if r._group is not None:
                raise RuntimeError("processChunks needs the same resamplers in the same order on every call "
                                   "(a resampler already belongs to another batch)")
        g = cls(len(resamplers), *key)
        # an instance whose `kernel` was set explicitly (e.g. pinned to STRICT by a test) keeps it
        # in the batch; otherwise the class-level batch default applies
        wanted = r0.__dict__.get("kernel", r0.batch_kernel)
        if wanted != KERNEL_AUTO:
            g.set_kernel(wanted)
        info = g.filter_info()
        hist = np.zeros(max((info.filt_len - 1) * g.channels, 1), dtype=np.int16)
        L = _lib.lib()
        for i, r in enumerate(resamplers):
            if r._resamplerPtr:
                # migrate a stream that already ran through processChunk: copy its device state
                st = _single_state(r, hist)
                L.spxb_batch_set_state(g._h, i, st[0], st[1], hist.ctypes.data)
                r.destroy()
            g._out_buffer_size[i] = r._outBufferSize
            r._group, r._group_index = g, i
        g.members = tuple(resamplers)
        return g

    def close(self):
        if getattr(self, "_h", None):
            _lib.lib().spxb_batch_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- knobs -------------------------------------------------------------------
    def set_kernel(self, kernel: int):
        e = _lib.lib().spxb_batch_set_kernel(self._h, kernel)
        if e:
            raise ValueError(_lib.strerror(e))

    def last_kernel(self) -> int:
        return _lib.lib().spxb_batch_get_kernel(self._h)

    def tensor_geometry(self):
        """{nt, ksteps, tiles, groups, stages, smem_bytes} of the last tensor-kernel call, or None"""
        import numpy as np
        g = np.zeros(6, np.uint32)
        if _lib.lib().spxb_batch_tensor_geometry(self._h, g.ctypes.data) != 0:
            return None
        return dict(zip(("nt", "ksteps", "tiles", "groups", "stages", "smem_bytes"), (int(v) for v in g)))

    def filter_info(self) -> _lib.FilterInfo:
        info = _lib.FilterInfo()
        _lib.lib().spxb_filter_describe(int(self.inRate), int(self.outRate), int(self.quality), C.byref(info))
        return info

    def counters(self) -> _lib.Counters:
        c = _lib.Counters()
        _lib.lib().spxb_batch_counters(self._h, C.byref(c))
        return c

    def synchronize(self):
        e = _lib.lib().spxb_batch_synchronize(self._h)
        if e:
            raise RuntimeError(_lib.strerror(e) + ": " + _lib.last_error())

    def get_state(self, stream: int):
        info = self.filter_info()
        hist = np.zeros(max((info.filt_len - 1) * self.channels, 1), dtype=np.int16)
        ls, fr, mg = C.c_int32(), C.c_uint32(), C.c_uint32()
        e = _lib.lib().spxb_batch_get_state(self._h, stream, C.byref(ls), C.byref(fr), C.byref(mg),
                                            hist.ctypes.data)
        if e:
            raise RuntimeError(_lib.strerror(e) + ": " + _lib.last_error())
        return ls.value, fr.value, mg.value, hist[: (info.filt_len - 1) * self.channels]

    def set_state(self, stream: int, last_sample: int, samp_frac_num: int, history: np.ndarray):
        h = np.ascontiguousarray(history, dtype=np.int16)
        e = _lib.lib().spxb_batch_set_state(self._h, stream, last_sample, samp_frac_num, h.ctypes.data)
        if e:
            raise RuntimeError(_lib.strerror(e) + ": " + _lib.last_error())

    def reset(self):
        _lib.lib().spxb_batch_reset(self._h)

    # -- processing ----------------------------------------------------------------
    def process(self, pcm: np.ndarray, in_frames, out_cap):
        """Speex-convention batched call on a packed array.

        pcm: int16 [n_streams, stride_frames*channels]; in_frames / out_cap: scalars or
        [n_streams]. Returns (out [n_streams, max_cap*channels], consumed[n], written[n])."""
        L = _lib.lib()
        pcm = np.ascontiguousarray(pcm, dtype=np.int16).reshape(self.n_streams, -1)
        in_stride = pcm.shape[1] // self.channels
        nin = np.ascontiguousarray(np.broadcast_to(np.asarray(in_frames, dtype=np.uint32), (self.n_streams,)))
        nout = np.ascontiguousarray(np.broadcast_to(np.asarray(out_cap, dtype=np.uint32), (self.n_streams,)))
        nin, nout = nin.copy(), nout.copy()
        out_stride = max(int(nout.max()), 1)
        out = np.zeros((self.n_streams, out_stride * self.channels), dtype=np.int16)
        e = L.spxb_batch_process(self._h, pcm.ctypes.data, in_stride, nin.ctypes.data,
                                 out.ctypes.data, out_stride, nout.ctypes.data)
        if e:
            raise RuntimeError(_lib.strerror(e) + ": " + _lib.last_error())
        return out, nin, nout

    def process_f32(self, pcm: np.ndarray, in_frames, out_cap):
        """speex_resampler_process_interleaved_float for every stream (float batch only):
        pcm float32 [n_streams, stride_frames*channels]; results are the kernels' f32 values,
        unrounded. Returns (out, consumed[n], written[n]) like process()."""
        L = _lib.lib()
        pcm = np.ascontiguousarray(pcm, dtype=np.float32).reshape(self.n_streams, -1)
        in_stride = pcm.shape[1] // self.channels
        nin = np.ascontiguousarray(np.broadcast_to(np.asarray(in_frames, dtype=np.uint32), (self.n_streams,))).copy()
        nout = np.ascontiguousarray(np.broadcast_to(np.asarray(out_cap, dtype=np.uint32), (self.n_streams,))).copy()
        out_stride = max(int(nout.max()), 1)
        out = np.zeros((self.n_streams, out_stride * self.channels), dtype=np.float32)
        e = L.spxb_batch_process_f32(self._h, pcm.ctypes.data, in_stride, nin.ctypes.data,
                                     out.ctypes.data, out_stride, nout.ctypes.data)
        if e:
            raise RuntimeError(_lib.strerror(e) + ": " + _lib.last_error())
        return out, nin, nout

    def process_pcm_f32(self, pcm: np.ndarray, in_frames, out_cap):
        """Scaled float PCM (+-1.0 full scale) in and out on an int16 batch: converted to int16 on load
        and back on store inside the kernel. Shapes and results like process()."""
        L = _lib.lib()
        pcm = np.ascontiguousarray(pcm, dtype=np.float32).reshape(self.n_streams, -1)
        in_stride = pcm.shape[1] // self.channels
        nin = np.ascontiguousarray(np.broadcast_to(np.asarray(in_frames, dtype=np.uint32), (self.n_streams,))).copy()
        nout = np.ascontiguousarray(np.broadcast_to(np.asarray(out_cap, dtype=np.uint32), (self.n_streams,))).copy()
        out_stride = max(int(nout.max()), 1)
        out = np.zeros((self.n_streams, out_stride * self.channels), dtype=np.float32)
        e = L.spxb_batch_process_pcm_f32(self._h, pcm.ctypes.data, in_stride, nin.ctypes.data,
                                         out.ctypes.data, out_stride, nout.ctypes.data)
        if e:
            raise RuntimeError(_lib.strerror(e) + ": " + _lib.last_error())
        return out, nin, nout

    def processChunks(self, chunks: Sequence) -> List[bytes]:
        """One processChunk per stream (the capacity rule of index.ts:80-95 per stream)."""
        if len(chunks) != self.n_streams:
            raise ValueError("processChunks needs one chunk per stream")
        datas = [_chunk_bytes(c) for c in chunks]
        frame_bytes = self.channels * 2
        lens = np.fromiter((len(d) for d in datas), dtype=np.int64, count=self.n_streams)
        if np.any(lens % frame_bytes != 0):
            raise ValueError("Chunk length should be a multiple of channels * 2 bytes")
        target = np.ceil(lens * float(self.outRate) / float(self.inRate))
        np.maximum(self._out_buffer_size, target, out=self._out_buffer_size)
        for i, r in enumerate(self.members):
            r._outBufferSize = self._out_buffer_size[i]
        caps = np.trunc(self._out_buffer_size / self.channels / 2).astype(np.uint32)
        frames = (lens // frame_bytes).astype(np.uint32)
        stride = max(int(frames.max()), 1)
        if int(lens.min()) == int(lens.max()):
            pcm = np.frombuffer(b"".join(datas), dtype=np.int16).reshape(self.n_streams, -1) \
                if lens[0] else np.zeros((self.n_streams, self.channels), np.int16)
        else:
            pcm = np.zeros((self.n_streams, stride * self.channels), dtype=np.int16)
            for i, d in enumerate(datas):
                pcm[i, : len(d) // 2] = np.frombuffer(d, dtype=np.int16)
        out, _, written = self.process(pcm, frames, caps)
        return [out[i, : int(written[i]) * self.channels].tobytes() for i in range(self.n_streams)]


def _single_state(r: SpeexResampler, hist: np.ndarray):
    """(last_sample, samp_frac_num) + history of a stream that lives on its own handle."""
    L = _lib.lib()
    batch_ptr = L.spxb_resampler_batch(r._resamplerPtr)
    ls, fr, mg = C.c_int32(), C.c_uint32(), C.c_uint32()
    e = L.spxb_batch_get_state(batch_ptr, 0, C.byref(ls), C.byref(fr), C.byref(mg), hist.ctypes.data)
    if e:
        raise RuntimeError(_lib.strerror(e) + ": " + _lib.last_error())
    return ls.value, fr.value


class SpeexResamplerTransform:
    """Stream adapter (src/index.ts:121-162): carries ``len % (channels*2)`` stray bytes to
    the next chunk and forwards aligned data to processChunk. ``_transform`` has the Node
    signature (chunk, encoding, callback); ``transform`` / ``pipe`` are Python conveniences.
    Like the reference there is no flush: trailing stray bytes are never emitted."""

    def __init__(self, channels, inRate, outRate, quality=7):
        self.resampler = SpeexResampler(channels, inRate, outRate, quality)
        self.channels, self.inRate, self.outRate, self.quality = channels, inRate, outRate, quality
        self._alignementBuffer = b""

    def _transform(self, chunk, encoding, callback):
        chunkToProcess = _chunk_bytes(chunk)
        if len(self._alignementBuffer) > 0:
            chunkToProcess = self._alignementBuffer + chunkToProcess
            self._alignementBuffer = b""
        extraneous = len(chunkToProcess) % (self.channels * 2)
        if extraneous != 0:
            self._alignementBuffer = chunkToProcess[len(chunkToProcess) - extraneous:]
            chunkToProcess = chunkToProcess[: len(chunkToProcess) - extraneous]
        try:
            res = self.resampler.processChunk(chunkToProcess)
        except Exception as e:  # index.ts:157-159 forwards the error to the callback
            callback(e, None)
            return
        callback(None, res)

    def transform(self, chunk) -> bytes:
        box = {}

        def cb(err, res):
            box["err"], box["res"] = err, res
        self._transform(chunk, None, cb)
        if box["err"] is not None:
            raise box["err"]
        return box["res"]

    def pipe(self, chunks: Iterable) -> Iterable[bytes]:
        for c in chunks:
            yield self.transform(c)


class SpeexResamplerBatchTransform:
    """Multi-stream counterpart of SpeexResamplerTransform: every written
    item is a sequence with one chunk per stream (any byte lengths, possibly empty), every
    result the list of resampled chunks. Each stream keeps its own alignment carry
    (src/index.ts:139-154 per stream); all streams of a write go to the GPU in one launch
    through SpeexResampler.processChunks, so stream i's output equals what its own
    SpeexResamplerTransform would have produced from the same writes."""

    def __init__(self, streams, channels, inRate, outRate, quality=7):
        self.streams, self.channels = int(streams), channels
        self.inRate, self.outRate, self.quality = inRate, outRate, quality
        self.resamplers = [SpeexResampler(channels, inRate, outRate, quality) for _ in range(self.streams)]
        self._alignementBuffers = [b""] * self.streams

    def _transform(self, chunks, encoding, callback):
        if len(chunks) != self.streams:
            callback(ValueError("SpeexResamplerBatchTransform expects one chunk per stream"), None)
            return
        frame_bytes = self.channels * 2
        whole = []
        for i, c in enumerate(chunks):
            data = self._alignementBuffers[i] + _chunk_bytes(c)
            extraneous = len(data) % frame_bytes
            self._alignementBuffers[i] = data[len(data) - extraneous:] if extraneous else b""
            whole.append(data[: len(data) - extraneous] if extraneous else data)
        try:
            res = SpeexResampler.processChunks(self.resamplers, whole)
        except Exception as e:
            callback(e, None)
            return
        callback(None, res)

    def transform(self, chunks) -> List[bytes]:
        box = {}

        def cb(err, res):
            box["err"], box["res"] = err, res
        self._transform(chunks, None, cb)
        if box["err"] is not None:
            raise box["err"]
        return box["res"]
