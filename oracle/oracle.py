"""ctypes drivers for the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Three checkers live here:

* ``OracleResampler``  -> oracle/liboracle.so, our C restatement (speex_oracle.c)
* ``RefResampler``     -> oracle/_ref/libspeex_ref.so, the reference's own
  deps/speex/resample.c compiled natively by oracle/Makefile (exists only where
  that build ran; it travels to the GPU box as a prebuilt, git-ignored .so)
* ``WasmResampler``    -> oracle/_ref/libspeex_wasm.so, the reference's SHIPPED WebAssembly
  module (embedded in src/speex_wasm.js), translated to C by oracle/wasm2c_lite.py and driven
  through its own malloc / length cells like src/index.ts does; same build and travel rules

All expose the reference wrapper's ``processChunk`` rule (src/index.ts:50-116)
so tests read like src/test.ts. Nothing in node_speex_resampler_b200/ imports
this module; only tests/, __graft_entry__.smoke() and bench.py's CPU legs do.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(HERE, "liboracle.so")
REF_SO = os.path.join(HERE, "_ref", "libspeex_ref.so")
# the reference's own test inputs: oracle/Makefile copies resources/*.pcm beside the built reference
# objects (git-ignored, travels to the GPU box); in the build container the originals are there too
FIXTURE_DIRS = (os.path.join(HERE, "_ref", "resources"), "/root/reference/resources")


def fixture_path(name: str):
    """Path of one of the reference's resources/*.pcm files, or None where neither copy exists."""
    for d in FIXTURE_DIRS:
        p = os.path.join(d, name)
        if os.path.isfile(p):
            return p
    return None


def build(quiet: bool = True) -> None:
    """Compile liboracle.so (always) and _ref/libspeex_ref.so + _ref/libspeex_wasm.so (when
    /root/reference exists)."""
    subprocess.run(["make", "-C", HERE] + (["-s"] if quiet else []), check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def have_ref() -> bool:
    return os.path.exists(REF_SO)


class OrcParams(C.Structure):
    _fields_ = [("num", C.c_uint32), ("den", C.c_uint32), ("filt_len", C.c_uint32),
                ("oversample", C.c_uint32), ("int_advance", C.c_int32),
                ("frac_advance", C.c_int32), ("cutoff", C.c_float),
                ("use_direct", C.c_int32), ("use_double", C.c_int32),
                ("table_len", C.c_uint32), ("channels", C.c_uint32),
                ("quality", C.c_int32)]


_orc = None
_ref = None


def _load_oracle():
    global _orc
    if _orc is None:
        if not os.path.exists(ORACLE_SO):
            build()
        L = C.CDLL(ORACLE_SO)
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(C.c_int)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_process_interleaved_int16.restype = C.c_int
        L.orc_process_interleaved_int16.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32),
                                                    C.c_void_p, C.POINTER(C.c_uint32)]
        L.orc_process_interleaved_float.restype = C.c_int
        L.orc_process_interleaved_float.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32),
                                                    C.c_void_p, C.POINTER(C.c_uint32)]
        L.orc_get_params.argtypes = [C.c_void_p, C.POINTER(OrcParams)]
        L.orc_table.restype = C.POINTER(C.c_float)
        L.orc_table.argtypes = [C.c_void_p]
        L.orc_get_state.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_int32),
                                    C.POINTER(C.c_uint32), C.POINTER(C.POINTER(C.c_float))]
        L.orc_process_chunk.restype = C.c_long
        L.orc_process_chunk.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_uint32, C.c_uint32,
                                        C.c_void_p, C.c_size_t, C.c_void_p]
        _orc = L
    return _orc


def _load_ref():
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            raise FileNotFoundError(f"{REF_SO} not built (reference tree absent?)")
        L = C.CDLL(REF_SO)
        L.speex_resampler_init.restype = C.c_void_p
        L.speex_resampler_init.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int,
                                           C.POINTER(C.c_int)]
        L.speex_resampler_destroy.argtypes = [C.c_void_p]
        L.speex_resampler_process_interleaved_int.restype = C.c_int
        L.speex_resampler_process_interleaved_int.argtypes = [
            C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32), C.c_void_p, C.POINTER(C.c_uint32)]
        L.speex_resampler_process_interleaved_float.restype = C.c_int
        L.speex_resampler_process_interleaved_float.argtypes = [
            C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32), C.c_void_p, C.POINTER(C.c_uint32)]
        L.speex_resampler_strerror.restype = C.c_char_p
        L.speex_resampler_strerror.argtypes = [C.c_int]
        L.speex_resampler_get_rate.argtypes = [C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        _ref = L
    return _ref


class _RefState(C.Structure):
    """LP64 layout of struct SpeexResamplerState_ (resample.c:116-146); test-only peek."""
    _fields_ = [("in_rate", C.c_uint32), ("out_rate", C.c_uint32), ("num_rate", C.c_uint32),
                ("den_rate", C.c_uint32), ("quality", C.c_int), ("nb_channels", C.c_uint32),
                ("filt_len", C.c_uint32), ("mem_alloc_size", C.c_uint32),
                ("buffer_size", C.c_uint32), ("int_advance", C.c_int), ("frac_advance", C.c_int),
                ("cutoff", C.c_float), ("oversample", C.c_uint32), ("initialised", C.c_int),
                ("started", C.c_int), ("last_sample", C.POINTER(C.c_int32)),
                ("samp_frac_num", C.POINTER(C.c_uint32)), ("magic_samples", C.POINTER(C.c_uint32)),
                ("mem", C.POINTER(C.c_float)), ("sinc_table", C.POINTER(C.c_float)),
                ("sinc_table_length", C.c_uint32), ("resampler_ptr", C.c_void_p),
                ("in_stride", C.c_int), ("out_stride", C.c_int)]


def _as_i16(buf) -> np.ndarray:
    a = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf
    if a.dtype != np.int16:
        a = np.ascontiguousarray(a).view(np.uint8)
    return a


class _ChunkRule:
    """The processChunk output-capacity rule of src/index.ts:80-95, shared by both drivers."""

    def __init__(self, channels, in_rate, out_rate):
        self.channels, self.in_rate, self.out_rate = channels, in_rate, out_rate
        self._out_buffer_size = -1.0

    def capacity_frames(self, nbytes: int) -> int:
        target = math.ceil(nbytes * self.out_rate / self.in_rate)
        if self._out_buffer_size < target:
            self._out_buffer_size = target
        return int(self._out_buffer_size / self.channels / 2)  # setValue(...,'i32') truncates


class OracleResampler:
    """Our C restatement behind the reference wrapper's surface."""

    def __init__(self, channels, in_rate, out_rate, quality=7):
        self.L = _load_oracle()
        err = C.c_int(0)
        self.h = self.L.orc_create(channels, in_rate, out_rate, quality, C.byref(err))
        if not self.h:
            raise ValueError(f"orc_create failed: err={err.value}")
        self.channels, self.in_rate, self.out_rate, self.quality = channels, in_rate, out_rate, quality
        self._rule = _ChunkRule(channels, in_rate, out_rate)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_destroy(self.h)
            self.h = None

    @property
    def params(self) -> OrcParams:
        p = OrcParams()
        self.L.orc_get_params(self.h, C.byref(p))
        return p

    def table(self) -> np.ndarray:
        p = self.params
        return np.ctypeslib.as_array(self.L.orc_table(self.h), shape=(p.table_len,)).copy()

    def state(self, channel=0):
        ls, fr, hp = C.c_int32(), C.c_uint32(), C.POINTER(C.c_float)()
        self.L.orc_get_state(self.h, channel, C.byref(ls), C.byref(fr), C.byref(hp))
        n = self.params.filt_len - 1
        hist = np.ctypeslib.as_array(hp, shape=(n,)).copy() if n else np.zeros(0, np.float32)
        return ls.value, fr.value, hist

    def process(self, pcm: np.ndarray, out_cap_frames: int):
        """speex_resampler_process_interleaved_int contract. pcm: int16 [frames*channels]."""
        pcm = np.ascontiguousarray(pcm, dtype=np.int16).reshape(-1)
        n_in = C.c_uint32(pcm.size // self.channels)
        n_out = C.c_uint32(out_cap_frames)
        out = np.empty(max(1, out_cap_frames * self.channels), dtype=np.int16)
        e = self.L.orc_process_interleaved_int16(self.h, pcm.ctypes.data, C.byref(n_in),
                                                 out.ctypes.data, C.byref(n_out))
        assert e == 0, e
        return out[: n_out.value * self.channels].copy(), n_in.value, n_out.value

    def process_float(self, pcm: np.ndarray, out_cap_frames: int):
        """speex_resampler_process_interleaved_float contract (float build). pcm: float32."""
        pcm = np.ascontiguousarray(pcm, dtype=np.float32).reshape(-1)
        n_in = C.c_uint32(pcm.size // self.channels)
        n_out = C.c_uint32(out_cap_frames)
        out = np.empty(max(1, out_cap_frames * self.channels), dtype=np.float32)
        e = self.L.orc_process_interleaved_float(self.h, pcm.ctypes.data, C.byref(n_in),
                                                 out.ctypes.data, C.byref(n_out))
        assert e == 0, e
        return out[: n_out.value * self.channels].copy(), n_in.value, n_out.value

    def processChunk(self, chunk) -> bytes:
        b = bytes(chunk) if not isinstance(chunk, np.ndarray) else chunk.tobytes()
        if len(b) % (self.channels * 2) != 0:
            raise ValueError("Chunk length should be a multiple of channels * 2 bytes")
        cap = self._rule.capacity_frames(len(b))
        out, _, _ = self.process(np.frombuffer(b, dtype=np.int16), cap)
        return out.tobytes()


class RefResampler:
    """The reference's own C (native gcc build) behind the same surface."""

    def __init__(self, channels, in_rate, out_rate, quality=7):
        self.L = _load_ref()
        err = C.c_int(0)
        self.h = self.L.speex_resampler_init(channels, in_rate, out_rate, quality, C.byref(err))
        if not self.h:
            raise ValueError(self.L.speex_resampler_strerror(err.value).decode())
        self.channels, self.in_rate, self.out_rate, self.quality = channels, in_rate, out_rate, quality
        self._rule = _ChunkRule(channels, in_rate, out_rate)

    def __del__(self):
        if getattr(self, "h", None):
            self.L.speex_resampler_destroy(self.h)
            self.h = None

    def _st(self) -> _RefState:
        return C.cast(self.h, C.POINTER(_RefState)).contents

    @property
    def params(self) -> OrcParams:
        s = self._st()
        p = OrcParams()
        p.num, p.den, p.filt_len, p.oversample = s.num_rate, s.den_rate, s.filt_len, s.oversample
        p.int_advance, p.frac_advance, p.cutoff = s.int_advance, s.frac_advance, s.cutoff
        p.table_len = (s.filt_len * s.den_rate if s.filt_len * s.den_rate <= s.filt_len * s.oversample + 8
                       else s.filt_len * s.oversample + 8)
        p.use_direct = int(s.filt_len * s.den_rate <= s.filt_len * s.oversample + 8)
        p.use_double = int(s.quality > 8)
        p.channels, p.quality = s.nb_channels, s.quality
        return p

    def table(self) -> np.ndarray:
        s = self._st()
        return np.ctypeslib.as_array(s.sinc_table, shape=(self.params.table_len,)).copy()

    def state(self, channel=0):
        s = self._st()
        n = s.filt_len - 1
        base = channel * s.mem_alloc_size
        hist = np.array([s.mem[base + j] for j in range(n)], dtype=np.float32)
        return s.last_sample[channel], s.samp_frac_num[channel], hist

    def process(self, pcm: np.ndarray, out_cap_frames: int):
        pcm = np.ascontiguousarray(pcm, dtype=np.int16).reshape(-1)
        n_in = C.c_uint32(pcm.size // self.channels)
        n_out = C.c_uint32(out_cap_frames)
        out = np.empty(max(1, out_cap_frames * self.channels), dtype=np.int16)
        e = self.L.speex_resampler_process_interleaved_int(
            self.h, pcm.ctypes.data, C.byref(n_in), out.ctypes.data, C.byref(n_out))
        if e != 0:
            raise RuntimeError(self.L.speex_resampler_strerror(e).decode())
        return out[: n_out.value * self.channels].copy(), n_in.value, n_out.value

    def process_float(self, pcm: np.ndarray, out_cap_frames: int):
        pcm = np.ascontiguousarray(pcm, dtype=np.float32).reshape(-1)
        n_in = C.c_uint32(pcm.size // self.channels)
        n_out = C.c_uint32(out_cap_frames)
        out = np.empty(max(1, out_cap_frames * self.channels), dtype=np.float32)
        e = self.L.speex_resampler_process_interleaved_float(
            self.h, pcm.ctypes.data, C.byref(n_in), out.ctypes.data, C.byref(n_out))
        if e != 0:
            raise RuntimeError(self.L.speex_resampler_strerror(e).decode())
        return out[: n_out.value * self.channels].copy(), n_in.value, n_out.value

    def processChunk(self, chunk) -> bytes:
        b = bytes(chunk) if not isinstance(chunk, np.ndarray) else chunk.tobytes()
        if len(b) % (self.channels * 2) != 0:
            raise ValueError("Chunk length should be a multiple of channels * 2 bytes")
        cap = self._rule.capacity_frames(len(b))
        out, _, _ = self.process(np.frombuffer(b, dtype=np.int16), cap)
        return out.tobytes()


WASM_SO = os.path.join(HERE, "_ref", "libspeex_wasm.so")
_wasm = None


def have_wasm() -> bool:
    return os.path.exists(WASM_SO)


def _load_wasm():
    global _wasm
    if _wasm is None:
        if not os.path.exists(WASM_SO):
            raise FileNotFoundError(f"{WASM_SO} not built (reference tree absent?)")
        L = C.CDLL(WASM_SO)
        u32 = C.c_uint32
        L.wasm_memory.restype = C.POINTER(C.c_uint8)
        L.wasm_memory_bytes.restype = u32
        L.wasm_malloc.restype = u32
        L.wasm_malloc.argtypes = [u32]
        L.wasm_free.argtypes = [u32]
        L.wasm_speex_resampler_init.restype = u32
        L.wasm_speex_resampler_init.argtypes = [u32, u32, u32, u32, u32]
        L.wasm_speex_resampler_destroy.argtypes = [u32]
        L.wasm_speex_resampler_process_interleaved_int.restype = u32
        L.wasm_speex_resampler_process_interleaved_int.argtypes = [u32, u32, u32, u32, u32]
        L.wasm_speex_resampler_strerror.restype = u32
        L.wasm_speex_resampler_strerror.argtypes = [u32]
        _wasm = L
    return _wasm


class WasmResampler:
    """The reference's SHIPPED WebAssembly module (translated to C by oracle/wasm2c_lite.py,
    oracle/_ref/libspeex_wasm.so), driven exactly as src/index.ts:50-116 drives it: staging
    buffers malloc'd in linear memory, i32 length cells, the grow-only capacity rule."""

    def __init__(self, channels, in_rate, out_rate, quality=7):
        self.L = _load_wasm()
        self.channels, self.in_rate, self.out_rate, self.quality = channels, in_rate, out_rate, quality
        self._ptr = 0
        self._in_ptr = self._out_ptr = -1
        self._in_size = self._out_size = -1
        self._rule = _ChunkRule(channels, in_rate, out_rate)

    def __del__(self):
        try:
            for p in (self._in_ptr, self._out_ptr):
                if p != -1:
                    self.L.wasm_free(p)
            if self._ptr:
                self.L.wasm_speex_resampler_destroy(self._ptr)
                self._ptr = 0
        except Exception:
            pass

    def _mem(self):
        n = self.L.wasm_memory_bytes()
        return np.ctypeslib.as_array(self.L.wasm_memory(), shape=(n,))

    def _init(self):
        L = self.L
        err_ptr = L.wasm_malloc(4)
        self._ptr = L.wasm_speex_resampler_init(self.channels, self.in_rate, self.out_rate, self.quality, err_ptr)
        err = int(self._mem()[err_ptr:err_ptr + 4].view(np.int32)[0])
        if err != 0:
            self._ptr = 0
            raise RuntimeError(self.strerror(err))
        self._in_len_ptr, self._out_len_ptr = L.wasm_malloc(4), L.wasm_malloc(4)

    def strerror(self, err: int) -> str:
        p = self.L.wasm_speex_resampler_strerror(err)
        m = self._mem()
        q = p
        while m[q]:
            q += 1
        return bytes(m[p:q]).decode()

    def process(self, pcm: np.ndarray, out_cap_frames: int):
        """speex_resampler_process_interleaved_int inside the module's linear memory"""
        L = self.L
        if not self._ptr:
            self._init()
        pcm = np.ascontiguousarray(pcm, dtype=np.int16).reshape(-1)
        nbytes, obytes = pcm.size * 2, max(out_cap_frames * self.channels * 2, 2)
        if self._in_size < nbytes:
            if self._in_ptr != -1:
                L.wasm_free(self._in_ptr)
            self._in_ptr, self._in_size = L.wasm_malloc(max(nbytes, 2)), nbytes
        if self._out_size < obytes:
            if self._out_ptr != -1:
                L.wasm_free(self._out_ptr)
            self._out_ptr, self._out_size = L.wasm_malloc(obytes), obytes
        m = self._mem()
        m[self._in_ptr:self._in_ptr + nbytes] = pcm.view(np.uint8)
        m[self._in_len_ptr:self._in_len_ptr + 4].view(np.uint32)[0] = pcm.size // self.channels
        m[self._out_len_ptr:self._out_len_ptr + 4].view(np.uint32)[0] = out_cap_frames
        e = L.wasm_speex_resampler_process_interleaved_int(self._ptr, self._in_ptr, self._in_len_ptr, self._out_ptr,
                                                           self._out_len_ptr)
        if e != 0:
            raise RuntimeError(self.strerror(e))
        m = self._mem()
        used = int(m[self._in_len_ptr:self._in_len_ptr + 4].view(np.uint32)[0])
        made = int(m[self._out_len_ptr:self._out_len_ptr + 4].view(np.uint32)[0])
        out = m[self._out_ptr:self._out_ptr + made * self.channels * 2].view(np.int16).copy()
        return out, used, made

    def processChunk(self, chunk) -> bytes:
        b = bytes(chunk) if not isinstance(chunk, np.ndarray) else chunk.tobytes()
        if len(b) % (self.channels * 2) != 0:
            raise ValueError("Chunk length should be a multiple of channels * 2 bytes")
        cap = self._rule.capacity_frames(len(b))
        out, _, _ = self.process(np.frombuffer(b, dtype=np.int16), cap)
        return out.tobytes()

    def table(self) -> np.ndarray:
        """the module's sinc table (wasm32 struct: sinc_table @76, sinc_table_length @80)"""
        if not self._ptr:
            self._init()
        m = self._mem()
        p = int(m[self._ptr + 76:self._ptr + 80].view(np.uint32)[0])
        n = int(m[self._ptr + 80:self._ptr + 84].view(np.uint32)[0])
        return m[p:p + 4 * n].view(np.float32).copy()


def best_cpu_resampler():
    """Reference build when it exists on this machine, else the restatement."""
    return (RefResampler, "reference") if have_ref() else (OracleResampler, "port")


def fnv1a64(data) -> str:
    """FNV-1a-64 over the bytes, as hex (the hash SURVEY.md 8c quotes)."""
    L = _load_oracle()
    b = data.tobytes() if isinstance(data, np.ndarray) else bytes(data)
    L.orc_fnv1a64.restype = C.c_uint64
    L.orc_fnv1a64.argtypes = [C.c_char_p, C.c_size_t]
    return f"{L.orc_fnv1a64(b, len(b)):016x}"


def snr_db(ref: np.ndarray, got: np.ndarray) -> float:
    ref = ref.astype(np.float64)
    err = ref - got.astype(np.float64)
    n = float(np.sum(err * err))
    if n == 0.0:
        return float("inf")
    return 10.0 * math.log10(float(np.sum(ref * ref)) / n)
