// kernels_common.cuh -- device helpers shared by the strict and tiled kernels.
#pragma once

#include <cuda_runtime.h>

#include "device_types.h"

namespace spxb {

__device__ __forceinline__ StreamCall load_call(const CallArgs &a, uint32_t s) {
  return a.per_stream ? a.per_stream[s] : a.uniform;
}

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
  return a.in[static_cast<size_t>(s) * a.in_stride + static_cast<size_t>(f) * a.channels + c];
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
  if (src < hist_elems)
    v = a.hist_src[static_cast<size_t>(s) * a.hist_stride + src];
  else
    v = a.in[static_cast<size_t>(s) * a.in_stride + (src - hist_elems)];
  a.hist_dst[static_cast<size_t>(s) * a.hist_stride + e] = v;
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
