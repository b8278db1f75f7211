// kernels_common.cuh -- device helpers shared by the strict and tiled kernels.
#pragma once

#include <cuda_runtime.h>

#include "device_types.h"

namespace spxb {

__device__ __forceinline__ StreamCall load_call(const CallArgs &a, uint32_t s) {
  return a.per_stream ? a.per_stream[s] : a.uniform;
}

// streams a launch covers, and the i-th of them (CallArgs::ids)
__device__ __forceinline__ uint32_t launch_rows(const CallArgs &a) { return a.ids ? a.n_ids : a.n_streams; }
__device__ __forceinline__ uint32_t launch_stream(const CallArgs &a, uint32_t i) { return a.ids ? a.ids[i] : i; }

// X~[f] of (stream s, channel c): history for f < 0, this call's input for 0 <= f < n_in,
// 0 beyond (only ever multiplied by zero taps or discarded).
__device__ __forceinline__ int fetch_sample(const CallArgs &a, uint32_t s, int f, uint32_t c,
                                            uint32_t n_in) {
  if (f < 0) {
    const int hf = f + static_cast<int>(a.hist_frames);
    if (hf < 0) return 0;
    return a.hist_src[static_cast<size_t>(s) * a.hist_stride + static_cast<size_t>(hf) * a.channels + c];
  }
  if (static_cast<uint32_t>(f) >= n_in) return 0;
  return a.in[static_cast<size_t>(s) * a.in_stride + static_cast<size_t>(f) * a.in_step + c];
}

// scaled float PCM ([-1, 1) full scale) -> the int16 sample the path computes with: round to nearest
// even of x * 32768, saturated (CallArgs::fmt == 3)
__device__ __forceinline__ int16_t pcm16_from_float(float x) {
  return static_cast<int16_t>(max(-32768, min(32767, __float2int_rn(__fmul_rn(x, 32768.f)))));
}

// Format-generic variants (CallArgs::fmt): the sample as the f32 the reference holds in `mem`
// (resample.c:1005 converts int16 input exactly; the float entry stores its input as is).
template <int FMT>
__device__ __forceinline__ float fetch_sample_f(const CallArgs &a, uint32_t s, int f, uint32_t c, uint32_t n_in) {
  if (FMT == 0) return static_cast<float>(fetch_sample(a, s, f, c, n_in));
  if (FMT == 3) {  // int16 history, scaled float in/out
    if (f < 0) return static_cast<float>(fetch_sample(a, s, f, c, n_in));
    if (static_cast<uint32_t>(f) >= n_in) return 0.f;
    const float x = reinterpret_cast<const float *>(a.in + static_cast<size_t>(s) * a.in_stride)[static_cast<size_t>(f) * a.in_step + c];
    return static_cast<float>(pcm16_from_float(x));
  }
  if (f < 0) {
    const int hf = f + static_cast<int>(a.hist_frames);
    if (hf < 0) return 0.f;
    const float *h = reinterpret_cast<const float *>(a.hist_src + static_cast<size_t>(s) * a.hist_stride);
    return h[static_cast<size_t>(hf) * a.channels + c];
  }
  if (static_cast<uint32_t>(f) >= n_in) return 0.f;
  const size_t e = static_cast<size_t>(f) * a.in_step + c;
  if (FMT == 2) return reinterpret_cast<const float *>(a.in + static_cast<size_t>(s) * a.in_stride)[e];
  return static_cast<float>(a.in[static_cast<size_t>(s) * a.in_stride + e]);
}

// 16 bytes of one stream's PCM starting at frame f (CH == 2: 4 frames, CH == 1: 8 frames):
// history for f < 0, the call's input for f >= 0, zeros outside both.
// in_align: largest of 16 / 8 / 4 / 2 bytes that every input row start is aligned to.
template <int CH>
__device__ __forceinline__ uint4 fetch_raw16(const CallArgs &a, const StreamCall &sc, uint32_t s, int f,
                                             int in_align) {
  constexpr int FPI = 8 / CH;  // frames per item
  uint4 raw = make_uint4(0u, 0u, 0u, 0u);
  if (s >= a.n_streams) return raw;
  if (f < 0) {
    const int hf = f + static_cast<int>(a.hist_frames);
    if (hf >= 0)
      raw = __ldg(reinterpret_cast<const uint4 *>(a.hist_src + static_cast<size_t>(s) * a.hist_stride +
                                                  static_cast<size_t>(hf) * CH));
    return raw;
  }
  if (static_cast<uint32_t>(f) >= sc.n_in) return raw;
  const int16_t *src = a.in + static_cast<size_t>(s) * a.in_stride + static_cast<size_t>(f) * CH;
  const int avail = min(FPI, static_cast<int>(sc.n_in) - f);
  if (avail == FPI) {
    if (in_align == 16) return __ldg(reinterpret_cast<const uint4 *>(src));
    if (in_align == 8) {
      const uint2 lo = __ldg(reinterpret_cast<const uint2 *>(src));
      const uint2 hi = __ldg(reinterpret_cast<const uint2 *>(src) + 1);
      return make_uint4(lo.x, lo.y, hi.x, hi.y);
    }
    if (in_align == 4) {
      const uint32_t *p = reinterpret_cast<const uint32_t *>(src);
      return make_uint4(__ldg(p), __ldg(p + 1), __ldg(p + 2), __ldg(p + 3));
    }
  }
  // tail of the input, or rows that are only 2-byte aligned: sample by sample
  uint32_t w[4] = {0u, 0u, 0u, 0u};
  const int n = avail * CH;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (i < n) w[i >> 1] |= static_cast<uint32_t>(static_cast<uint16_t>(src[i])) << (16 * (i & 1));
  return make_uint4(w[0], w[1], w[2], w[3]);
}

// float -> int16 exactly as the reference's WORD2INT (deps/speex/arch.h:208-209):
// saturate at -32767.5 / +32766.5, otherwise floor(0.5 + x) evaluated in f64.
__device__ __forceinline__ int16_t word2int_exact(float v) {
  if (v < -32767.5f) return static_cast<int16_t>(-32768);
  if (v > 32766.5f) return static_cast<int16_t>(32767);
  return static_cast<int16_t>(__double2int_rd(__dadd_rn(0.5, static_cast<double>(v))));
}

// same mapping with the half-up rounding done in f32: differs from the exact form only when
// v lies within one f32 ulp below a .5 boundary (tiled kernel; inside its 1-LSB contract)
__device__ __forceinline__ int word2int_fast(float v) {
  int r = __float2int_rd(v + 0.5f);
  return max(-32768, min(32767, r));
}

// History slide fused into the FIR kernels (deps/speex/resample.c:898-899): element e of
// the new history = element consumed*channels + e of (old history || input).
__device__ __forceinline__ void slide_history_elem(const CallArgs &a, uint32_t s,
                                                   const StreamCall &sc, uint32_t e) {
  const uint32_t hist_elems = a.hist_frames * a.channels;
  const size_t src = static_cast<size_t>(sc.consumed) * a.channels + e;
  int16_t v;
  if (src < hist_elems) {
    v = a.hist_src[static_cast<size_t>(s) * a.hist_stride + src];
  } else {
    const size_t i = src - hist_elems;  // element of the call's input, frames channels apart unless strided
    v = a.in[static_cast<size_t>(s) * a.in_stride + (i / a.channels) * a.in_step + i % a.channels];
  }
  a.hist_dst[static_cast<size_t>(s) * a.hist_stride + e] = v;
}

template <int FMT>
__device__ __forceinline__ void slide_history_elem_f(const CallArgs &a, uint32_t s, const StreamCall &sc, uint32_t e) {
  if (FMT == 0) {
    slide_history_elem(a, s, sc, e);
    return;
  }
  if (FMT == 3) {  // int16 history fed from scaled float input
    const uint32_t he = a.hist_frames * a.channels;
    const size_t from = static_cast<size_t>(sc.consumed) * a.channels + e;
    int16_t v;
    if (from < he) {
      v = a.hist_src[static_cast<size_t>(s) * a.hist_stride + from];
    } else {
      const size_t i = from - he;
      v = pcm16_from_float(reinterpret_cast<const float *>(a.in + static_cast<size_t>(s) * a.in_stride)[(i / a.channels) * a.in_step + i % a.channels]);
    }
    a.hist_dst[static_cast<size_t>(s) * a.hist_stride + e] = v;
    return;
  }
  const uint32_t hist_elems = a.hist_frames * a.channels;
  const size_t src = static_cast<size_t>(sc.consumed) * a.channels + e;
  float v;
  const size_t i = src < hist_elems ? 0 : src - hist_elems;
  const size_t ie = (i / a.channels) * a.in_step + i % a.channels;  // frames in_step apart (== channels unless strided)
  if (src < hist_elems)
    v = reinterpret_cast<const float *>(a.hist_src + static_cast<size_t>(s) * a.hist_stride)[src];
  else if (FMT == 2)
    v = reinterpret_cast<const float *>(a.in + static_cast<size_t>(s) * a.in_stride)[ie];
  else
    v = static_cast<float>(a.in[static_cast<size_t>(s) * a.in_stride + ie]);
  reinterpret_cast<float *>(a.hist_dst + static_cast<size_t>(s) * a.hist_stride)[e] = v;
}

template <int FMT>
__device__ __forceinline__ void history_block_f(const CallArgs &a, uint32_t blk) {
  const uint32_t hist_elems = a.hist_frames * a.channels;
  const uint32_t per_stream = (hist_elems + blockDim.x - 1) / blockDim.x;
  const uint32_t row = blk / per_stream;
  if (row >= launch_rows(a)) return;
  const uint32_t s = launch_stream(a, row);
  const uint32_t e = (blk % per_stream) * blockDim.x + threadIdx.x;
  const StreamCall sc = load_call(a, s);
  if (e < hist_elems) slide_history_elem_f<FMT>(a, s, sc, e);
  if (e == 0) {
    a.last_sample[s] = sc.ls1;
    a.samp_frac[s] = sc.frac1;
  }
}

// Blocks appended to every FIR grid: slide the history of all streams and publish the new
// (last_sample, samp_frac_num). `blk` counts from 0 over hist_blocks(a) blocks.
__device__ __forceinline__ void history_block(const CallArgs &a, uint32_t blk) {
  const uint32_t hist_elems = a.hist_frames * a.channels;
  const uint32_t per_stream = (hist_elems + blockDim.x - 1) / blockDim.x;
  const uint32_t s = blk / per_stream;
  if (s >= a.n_streams) return;
  const uint32_t e = (blk % per_stream) * blockDim.x + threadIdx.x;
  const StreamCall sc = load_call(a, s);
  if (e < hist_elems) slide_history_elem(a, s, sc, e);
  if (e == 0) {
    a.last_sample[s] = sc.ls1;
    a.samp_frac[s] = sc.frac1;
  }
}

}  // namespace spxb
