// device_types.h -- plain structs shared by the host runtime and the CUDA kernels.
//
// HBM layout of one batch (all streams share channels / ratio / quality):
//   table       f32 [table_len]            reference-layout sinc table (strict kernel)
//   phase_taps  f32 [den][N]               one FIR per output phase   (tiled kernel)
//   blend       f32 [den][4]               cubic weights per phase    (strict, interpolate)
//   hist[2]     i16 [stream][hist_stride]  ping-pong history: frames f in [-hist_frames, 0)
//                                          interleaved [frame][channel]; the newest N-1 frames
//                                          are live, the leading pad (hist_frames is N-1
//                                          rounded up to 4) keeps 4-frame groups aligned
//   last_sample i32 [stream], samp_frac u32 [stream], magic u32 [stream]
// A call sees, per stream, X~[f] = hist for f < 0 and the call's input for f >= 0. Output m
// reads X~[q(m) .. q(m)+N-1] with q(m) = last_sample - (N-1) + floor((frac + m*num)/den)
// and phase (frac + m*num) % den  (deps/speex/resample.c:344-378 in closed form).
#pragma once

#include <cstddef>
#include <cstdint>

namespace spxb {

// one stream's share of a call; computed on the host by plan_call()
struct StreamCall {
  int32_t ls0;        // last_sample at entry
  uint32_t frac0;     // samp_frac_num at entry
  uint32_t n_in;      // input frames offered
  uint32_t n_out;     // output frames to write
  uint32_t consumed;  // input frames the call commits (history slides by this)
  int32_t ls1;        // state after the call
  uint32_t frac1;
  uint32_t pad_;
};

struct FilterDev {
  uint32_t num, den, taps, oversample;
  int32_t direct;      // reference table is per-phase
  int32_t wide_accum;  // reference accumulates in f64 (quality > 8)
  const float *table;
  const float *phase_taps;
  const float *blend;  // [den][4], interpolate path only
  // pre-shifted tap tiles for the streaming kernel (filter_bank.h: BandTable); nullptr when
  // the table would be too large for this ratio
  const float *band;
  uint32_t band_kp, band_pad, band_row;
};

struct CallArgs {
  FilterDev filt;
  uint32_t n_streams;
  uint32_t channels;
  const int16_t *in;    // device; stream s at in + s*in_stride (int16 elements)
  size_t in_stride;
  int16_t *out;         // device; stream s at out + s*out_stride
  size_t out_stride;
  const int16_t *hist_src;
  int16_t *hist_dst;
  uint32_t hist_stride;  // int16 elements per stream
  uint32_t hist_frames;  // frames stored per stream (N-1 rounded up to a multiple of 4)
  int32_t *last_sample;  // device state arrays, rewritten by the call
  uint32_t *samp_frac;
  const StreamCall *per_stream;  // device array [n_streams], or nullptr when uniform
  StreamCall uniform;            // used when per_stream == nullptr
  uint32_t max_n_out;            // max over streams (grid sizing)
  // Sample formats (strict kernel only beyond 0). The pointers above stay int16-typed and every
  // stride stays in int16 units; a float sample simply occupies two of them.
  //   0  int16 history, int16 in/out                 (the hot path)
  //   1  float history, int16 in/out                 (a state that has seen float calls)
  //   2  float history, float in/out, no rounding    (speex_resampler_process_interleaved_float)
  //   3  int16 history, SCALED float in/out          (float PCM, +-1.0 full scale: converted to the
  //      int16 sample on load and back on store, the int16 path in between -- spxb_batch_process_pcm_f32)
  uint32_t fmt;
  // Optional subset: when ids != nullptr the launch covers streams ids[0 .. n_ids) only (device
  // array), in that order, instead of 0 .. n_streams. Used to run the groups of a ragged batch
  // that share one position on the tensor kernel, and the remainder on the strict kernel.
  const uint32_t *ids;
  uint32_t n_ids;
  // Element distance between consecutive frames of a series in `in` / `out` (strict kernel only;
  // the fast kernels take channels-interleaved rows, i.e. step == channels). The per-channel
  // entries of the Speex API (resample.c:925-1036 with st->in_stride / st->out_stride) read and
  // write one channel at the caller's stride. Counted in samples of the call's format.
  uint32_t in_step, out_step;
};

}  // namespace spxb
