"""ctypes binding of libspeexb200.so (the C ABI declared in include/speexb200.h).

There is no fallback: if the CUDA library is missing this module raises at import, and if
no CUDA device is present every constructor fails with the library's own error text.
"""
from __future__ import annotations

import ctypes as C
import os

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPXB_LIB_PATH") or os.path.join(PKG_DIR, "libspeexb200.so")  # override: A/B builds

# symbols include/speexb200.h declares; tests assert the .so exports every one of them
DECLARED_SYMBOLS = (
    "speex_resampler_init", "speex_resampler_destroy", "speex_resampler_get_rate",
    "speex_resampler_process_interleaved_int", "speex_resampler_strerror",
    "speex_resampler_get_ratio", "speex_resampler_get_quality",
    "speex_resampler_get_input_latency", "speex_resampler_get_output_latency",
    "speex_resampler_skip_zeros", "speex_resampler_reset_mem",
    "spxb_device_count", "spxb_last_error", "spxb_batch_create", "spxb_batch_destroy",
    "spxb_batch_set_kernel", "spxb_batch_get_kernel", "spxb_batch_tensor_geometry", "spxb_batch_tensor_trace", "spxb_batch_process", "spxb_batch_submit",
    "spxb_batch_wait", "spxb_batch_pipeline_depth", "spxb_batch_process_device",
    "spxb_batch_process_device_uniform", "spxb_batch_process_device_ring", "spxb_batch_set_stream", "spxb_batch_use_own_stream", "spxb_batch_synchronize",
    "spxb_batch_get_state", "spxb_batch_set_state", "spxb_batch_reset", "spxb_batch_skip_zeros",
    "spxb_batch_counters", "spxb_host_alloc", "spxb_host_free", "spxb_filter_describe",
    "spxb_filter_table", "spxb_filter_phase_taps", "spxb_filter_fixed_taps", "spxb_tensor_plan",
    "spxb_tensor_tap_tile", "spxb_plan_call", "spxb_version", "spxb_resampler_batch",
    "speex_resampler_init_frac", "speex_resampler_set_rate", "speex_resampler_set_rate_frac",
    "speex_resampler_set_quality",
    "speex_resampler_process_interleaved_float", "spxb_batch_create_f32", "spxb_batch_is_f32",
    "spxb_batch_process_f32", "spxb_batch_get_state_f32", "spxb_batch_set_state_f32", "spxb_plan_call_f32", "spxb_plan_call_ex",
    "spxb_measure_fp32_peak", "spxb_tensor_packed_plan", "spxb_tensor_tap_tile_packed",
    "spxb_batch_process_strided", "spxb_batch_process_pcm_f32", "speex_resampler_process_int", "speex_resampler_process_float",
    "speex_resampler_set_input_stride", "speex_resampler_get_input_stride",
    "speex_resampler_set_output_stride", "speex_resampler_get_output_stride",
)

KERNEL_AUTO, KERNEL_STRICT, KERNEL_TILED, KERNEL_TENSOR = 0, 1, 2, 3


class FilterInfo(C.Structure):
    _fields_ = [("num", C.c_uint32), ("den", C.c_uint32), ("filt_len", C.c_uint32),
                ("oversample", C.c_uint32), ("int_advance", C.c_int32),
                ("frac_advance", C.c_int32), ("cutoff", C.c_float), ("use_direct", C.c_int32),
                ("use_double", C.c_int32), ("table_len", C.c_uint32)]


class CallPlan(C.Structure):
    _fields_ = [("n_out", C.c_uint32), ("consumed", C.c_uint32), ("last_sample", C.c_int32),
                ("samp_frac_num", C.c_uint32)]


class Counters(C.Structure):
    _fields_ = [("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64),
                ("d2h_bytes", C.c_uint64), ("calls", C.c_uint64)]


def _bind(L):
    u32, i32, vp, sz = C.c_uint32, C.c_int32, C.c_void_p, C.c_size_t
    pu32, pint = C.POINTER(C.c_uint32), C.POINTER(C.c_int)
    L.speex_resampler_init.restype = vp
    L.speex_resampler_init.argtypes = [u32, u32, u32, C.c_int, pint]
    L.speex_resampler_destroy.argtypes = [vp]
    L.speex_resampler_get_rate.argtypes = [vp, pu32, pu32]
    L.speex_resampler_get_ratio.argtypes = [vp, pu32, pu32]
    L.speex_resampler_get_quality.argtypes = [vp, pint]
    L.speex_resampler_process_interleaved_int.restype = C.c_int
    L.speex_resampler_process_interleaved_int.argtypes = [vp, vp, pu32, vp, pu32]
    L.speex_resampler_strerror.restype = C.c_char_p
    L.speex_resampler_strerror.argtypes = [C.c_int]
    for name in ("speex_resampler_get_input_latency", "speex_resampler_get_output_latency",
                 "speex_resampler_skip_zeros", "speex_resampler_reset_mem"):
        getattr(L, name).restype = C.c_int
        getattr(L, name).argtypes = [vp]
    L.spxb_device_count.restype = C.c_int
    L.spxb_last_error.restype = C.c_char_p
    L.spxb_version.restype = C.c_char_p
    L.spxb_batch_create.restype = vp
    L.spxb_batch_create.argtypes = [u32, u32, u32, u32, C.c_int, C.c_int, pint]
    L.spxb_batch_destroy.argtypes = [vp]
    L.spxb_batch_set_kernel.argtypes = [vp, C.c_int]
    L.spxb_batch_get_kernel.argtypes = [vp]
    L.spxb_batch_process.argtypes = [vp, vp, sz, vp, vp, sz, vp]
    L.spxb_batch_submit.argtypes = [vp, vp, sz, vp, vp, sz, vp, C.POINTER(C.c_uint64)]
    L.spxb_batch_wait.argtypes = [vp, C.c_uint64]
    L.spxb_batch_pipeline_depth.argtypes = [vp]
    L.spxb_batch_process_device.argtypes = [vp, vp, sz, vp, vp, sz, vp]
    L.spxb_batch_process_device_uniform.argtypes = [vp, vp, sz, u32, vp, sz, u32, pu32, pu32]
    L.spxb_batch_process_device_ring.argtypes = [vp, vp, sz, sz, vp, sz, sz, u32, u32, u32, u32, u32]
    L.spxb_batch_set_stream.argtypes = [vp, vp]
    L.spxb_batch_use_own_stream.argtypes = [vp]
    L.spxb_batch_synchronize.argtypes = [vp]
    L.spxb_batch_get_state.argtypes = [vp, u32, C.POINTER(i32), pu32, pu32, vp]
    L.spxb_batch_set_state.argtypes = [vp, u32, i32, u32, vp]
    L.spxb_batch_reset.argtypes = [vp]
    L.spxb_batch_skip_zeros.argtypes = [vp]
    L.spxb_batch_counters.argtypes = [vp, C.POINTER(Counters)]
    L.spxb_host_alloc.restype = vp
    L.spxb_host_alloc.argtypes = [sz]
    L.spxb_host_free.argtypes = [vp]
    L.spxb_filter_describe.argtypes = [u32, u32, C.c_int, C.POINTER(FilterInfo)]
    L.spxb_filter_table.restype = C.c_long
    L.spxb_filter_table.argtypes = [u32, u32, C.c_int, vp, sz]
    L.spxb_filter_phase_taps.restype = C.c_long
    L.spxb_filter_phase_taps.argtypes = [u32, u32, C.c_int, vp, sz]
    L.spxb_batch_tensor_geometry.argtypes = [vp, vp]
    L.spxb_batch_tensor_trace.restype = C.c_long
    L.spxb_batch_tensor_trace.argtypes = [vp, vp, sz]
    L.spxb_filter_fixed_taps.restype = C.c_long
    L.spxb_filter_fixed_taps.argtypes = [u32, u32, C.c_int, vp, sz, pint]
    L.spxb_tensor_plan.restype = C.c_long
    L.spxb_tensor_plan.argtypes = [u32, u32, C.c_int, i32, u32, u32, u32, vp, sz, pu32]
    L.spxb_tensor_tap_tile.restype = C.c_long
    L.spxb_tensor_tap_tile.argtypes = [u32, u32, C.c_int, u32, u32, u32, vp, sz]
    L.spxb_tensor_packed_plan.restype = C.c_long
    L.spxb_tensor_packed_plan.argtypes = [u32, u32, C.c_int, u32, vp, sz, pu32]
    L.spxb_tensor_tap_tile_packed.restype = C.c_long
    L.spxb_tensor_tap_tile_packed.argtypes = [u32, u32, C.c_int, u32, u32, u32, vp, sz]
    L.spxb_batch_process_pcm_f32.argtypes = [vp, vp, sz, vp, vp, sz, vp]
    L.spxb_batch_process_strided.argtypes = [vp, vp, sz, u32, vp, vp, sz, u32, vp, C.c_int]
    L.speex_resampler_process_int.argtypes = [vp, u32, vp, pu32, vp, pu32]
    L.speex_resampler_process_float.argtypes = [vp, u32, vp, pu32, vp, pu32]
    L.speex_resampler_set_input_stride.argtypes = [vp, u32]
    L.speex_resampler_set_output_stride.argtypes = [vp, u32]
    L.speex_resampler_get_input_stride.argtypes = [vp, pu32]
    L.speex_resampler_get_output_stride.argtypes = [vp, pu32]
    L.spxb_plan_call.argtypes = [u32, u32, i32, u32, u32, u32, C.POINTER(CallPlan)]
    L.spxb_plan_call_f32.argtypes = [u32, u32, i32, u32, u32, u32, C.POINTER(CallPlan)]
    L.spxb_plan_call_ex.argtypes = [u32, u32, i32, u32, u32, u32, u32, C.c_int, u32, C.POINTER(CallPlan), pu32]
    L.speex_resampler_init_frac.restype = vp
    L.speex_resampler_init_frac.argtypes = [u32, u32, u32, u32, u32, C.c_int, pint]
    L.speex_resampler_set_rate.argtypes = [vp, u32, u32]
    L.speex_resampler_set_rate_frac.argtypes = [vp, u32, u32, u32, u32]
    L.speex_resampler_set_quality.argtypes = [vp, C.c_int]
    L.speex_resampler_process_interleaved_float.restype = C.c_int
    L.speex_resampler_process_interleaved_float.argtypes = [vp, vp, pu32, vp, pu32]
    L.spxb_batch_create_f32.restype = vp
    L.spxb_batch_create_f32.argtypes = [u32, u32, u32, u32, C.c_int, C.c_int, pint]
    L.spxb_batch_is_f32.argtypes = [vp]
    L.spxb_batch_process_f32.argtypes = [vp, vp, sz, vp, vp, sz, vp]
    L.spxb_batch_get_state_f32.argtypes = [vp, u32, C.POINTER(i32), pu32, pu32, vp]
    L.spxb_batch_set_state_f32.argtypes = [vp, u32, i32, u32, vp]
    L.spxb_resampler_batch.restype = vp
    L.spxb_resampler_batch.argtypes = [vp]
    L.spxb_measure_fp32_peak.restype = C.c_double
    L.spxb_measure_fp32_peak.argtypes = [C.c_int]
    return L


_lib = None


def lib():
    """The loaded library. Raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (nvcc, sm_100a). node_speex_resampler_b200 has no CPU path.")
        _lib = _bind(C.CDLL(LIB_PATH))
    return _lib


def strerror(code: int) -> str:
    return lib().speex_resampler_strerror(int(code)).decode()


def last_error() -> str:
    return lib().spxb_last_error().decode()
